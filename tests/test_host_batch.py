"""Host side of the batch path (calfkit/engine/lane.py, calfkit/broker.py) without a GPU: the containers a lane hands to the
broker are numpy views over landing buffers — built here by hand in exactly the layout the engine fills (payloads at 16-byte
aligned starts, the ck_publish table, the device-side grouping by topic) — and everything the Worker and the broker do with
them is index arithmetic that must agree with the per-record view."""
import json

import numpy as np
import pytest

from calfkit.broker import MemoryBroker, Record, _fit
from calfkit.engine._lib import PUB_DTYPE
from calfkit.engine.lane import NO_PAYLOAD, Arena, PublishBatch, received_batch


def _layout(payloads: list[bytes]):
    lens = np.asarray([len(p) for p in payloads], dtype=np.uint32)
    off = np.zeros(len(payloads) + 1, dtype=np.int64)
    np.cumsum((lens.astype(np.int64) + 15) & ~15, out=off[1:])
    out = np.zeros(int(off[-1]), dtype=np.uint8)
    for i, p in enumerate(payloads):
        out[off[i]:off[i] + len(p)] = np.frombuffer(p, dtype=np.uint8)
    return out, off, lens


def _batch(with_groups: bool):
    """three source records; record 0 and 2 publish to registered topics (ids 0 / 1) with a key, record 1 to a topic the
    engine has no id for (named by a span of the source record) and, unkeyed, to topic 1; one unused publish slot"""
    recs = [b'{"corr":"id-0","topic":"alpha"}', b'{"corr":"id\\"1","topic":"client-reply-7"}', b'{"corr":"id-2","topic":"beta"}']
    src = Arena.pack(recs)
    payloads = [b"P0" * 9, b"P1" * 20, b"P2" * 8, b"P3" * 8]          # 18, 40, 16, 16 bytes: only the 16-byte ones pack back to back
    out, off, lens = _layout(payloads)
    pubs = np.zeros(6, dtype=PUB_DTYPE)
    pubs["payload"] = [0, 1, 1, NO_PAYLOAD, 2, 3]
    pubs["topic_id"] = [0, -1, 1, -1, 1, 1]
    pubs["record"] = [0, 1, 1, 1, 2, 2]
    pubs["has_key"] = [1, 1, 0, 0, 1, 1]
    pubs["partition"] = [3, 5, -1, -1, 2, 2]
    t = recs[1].index(b"client-reply-7")
    pubs["topic_off"][1], pubs["topic_len"][1], pubs["pad"][1] = t, len(b"client-reply-7"), 0xABCD
    key_spans = np.asarray([[9, 9, 9], [4, 5, 4]], dtype=np.uint32)      # offset / length of the corr value in each record
    status, action = np.zeros(3, dtype=np.uint32), np.ones(3, dtype=np.uint32)
    order = key_counts = None
    if with_groups:
        # what ck_group_publishes leaves: indices grouped by key 0 (no id), 1 + id, 4095 (unused), send order kept
        order = np.asarray([1, 0, 2, 4, 5, 3], dtype=np.uint32)
        key_counts = np.zeros(4096, dtype=np.uint32)
        key_counts[0], key_counts[1], key_counts[2], key_counts[4095] = 1, 1, 3, 1
    released = []
    b = PublishBatch(out, off, lens, pubs, {0: "alpha", 1: "beta"}, src, key_spans, status, action, on_release=lambda: released.append(1),
                     order=order, key_counts=key_counts)
    return b, recs, payloads, released


@pytest.mark.parametrize("with_groups", [False, True])
def test_publish_batch_views(with_groups):
    b, recs, payloads, released = _batch(with_groups)
    assert b.n_publishes == 5
    assert b.topic_counts() == {-1: 1, 0: 1, 1: 3}
    assert b.select(0).tolist() == [0] and b.select(1).tolist() == [2, 4, 5] and b.select(-1).tolist() == [1]
    assert [b.payload(j) for j in (0, 1, 2, 4, 5)] == [payloads[0], payloads[1], payloads[1], payloads[2], payloads[3]]
    assert b.key(0) == b"id-0" and b.key(1) == b'id"1' and b.key(2) is None and b.key(4) == b"id-2"      # the key is the UNescaped id
    assert b.topic_name(0) == "alpha" and b.topic_name(1) == "client-reply-7" and b.topic_name(5) == "beta"
    got = list(b.iter_records())
    assert [(t, k, p) for t, k, p, _j in got] == [("alpha", b"id-0", payloads[0]), ("client-reply-7", b'id"1', payloads[1]),
                                                   ("beta", None, payloads[1]), ("beta", b"id-2", payloads[2]), ("beta", b"id-2", payloads[3])]
    # an in-process hop: consecutive 16-byte-multiple payloads come back as a zero-copy view that keeps the batch alive ...
    a = b.gather(np.asarray([4, 5]))
    assert a.n == 2 and a.record(0) == payloads[2] and a.record(1) == payloads[3] and np.shares_memory(a.data, b.out)
    b.release()
    assert not released                    # ... until the arena lets go
    a.release()
    assert released == [1]
    # ... anything else is packed
    b2, _r, payloads2, _rel = _batch(with_groups)
    a2 = b2.gather(np.asarray([0, 2]))
    assert [a2.record(i) for i in range(a2.n)] == [payloads2[0], payloads2[1]] and not np.shares_memory(a2.data, b2.out)
    assert b2.gather(np.asarray([], dtype=np.int64)).n == 0


def test_received_batch_is_a_publish_batch():
    from calfkit.engine.exchange import _meta_dtype
    payloads = [b"x" * 33, b"y" * 16, b"z" * 5]
    out, off, lens = _layout(payloads)
    meta = np.zeros(3, dtype=_meta_dtype())
    meta["len"], meta["topic_id"], meta["partition"], meta["src_pub"] = lens, [1, 0, 1], [4, 4, 6], [10, 11, 12]
    released = []
    rb = received_batch(meta, out, {0: "alpha", 1: "beta"}, on_release=lambda: released.append(1))
    assert rb.n_publishes == 3 and rb.topic_counts() == {0: 1, 1: 2}
    assert [rb.payload(j) for j in rb.select(1)] == [payloads[0], payloads[2]] and rb.payload(int(rb.select(0)[0])) == payloads[1]
    assert rb.key(0) is None and rb.pubs["partition"].tolist() == [4, 4, 6]
    rb.release()
    assert released == [1]


def test_broker_routes_publish_batches_without_touching_records():
    b, recs, payloads, released = _batch(True)
    br = MemoryBroker()
    br.subscriber("beta", group_id="g")                 # a node of this worker consumes beta
    seen = []
    br.sink("alpha", lambda batch, idx: seen.append([batch.payload(int(j)) for j in idx]))
    br.produce_publishes(b)
    assert seen == [[payloads[0]]]
    assert br.dropped_unsubscribed == 1                  # client-reply-7: nobody here subscribes, counted, not kept
    assert br.produced == 5 and br.pending() == 3
    assert not released                                  # the fragment for beta holds the batch
    # beta's three payloads come back as arenas bounded by records and bytes
    a1 = br.poll_arena(("beta",), max_records=2, max_bytes=1 << 20)
    assert [a1.record(i) for i in range(a1.n)] == [payloads[1], payloads[2]]
    a2 = br.poll_arena(("beta",), max_records=8, max_bytes=1 << 20)
    assert [a2.record(i) for i in range(a2.n)] == [payloads[3]]
    assert br.poll_arena(("beta",), 8, 1 << 20) is None and br.pending() == 0
    a1.release(); a2.release()
    assert released == [1]
    # a forwarded batch for a topic without an id cannot be named here: loud, counted
    rb = PublishBatch(b.out, b.out_off, b.out_len, np.asarray([(0, -1, 0, 0, 0, 0, 1, 0)], dtype=PUB_DTYPE), {}, None, None, None, None)
    br.produce_publishes(rb)
    assert br.dropped_unsubscribed == 2


def test_arenas_and_poll_limits():
    recs = [b"a" * 10, b"b" * 200, b"c" * 30, b"d" * 5]
    a = Arena.pack(recs)
    assert a.n == 4 and a.nbytes == 245 and [a.record(i) for i in range(4)] == recs
    s = a.slice(1, 3)
    assert s.n == 2 and s.record(0) == recs[1] and s.offsets[0] == 0 and np.shares_memory(s.data, a.data)
    assert Arena.pack([]).n == 0
    # _fit: records and bytes, never fewer than one record
    off = a.offsets
    assert _fit(off, 10, None) == 4 and _fit(off, 2, None) == 2 and _fit(off, 10, 210) == 2 and _fit(off, 10, 5) == 1 and _fit(off, 10, 245) == 4
    br = MemoryBroker()
    br.produce_arena("t", a)
    h = br.poll_arena(("t",), max_records=10, max_bytes=215)
    assert [h.record(i) for i in range(h.n)] == recs[:2]
    rest = br.poll_arena(("t",), 10, None)
    assert [rest.record(i) for i in range(rest.n)] == recs[2:] and br.pending() == 0
    # per-record and batch-wise entries share a topic: the per-record consumer sees both
    br.produce(Record("t", b"r1", b"k", "c"))
    br.produce_arena("t", Arena.pack([b"r2", b"r3"]))
    got = br.poll_batch(("t",), 10)
    assert [r.value for r in got] == [b"r1", b"r2", b"r3"] and got[0].key == b"k" and got[1].key is None
    # an oversized record is still handed out alone (the engine reports it per record)
    br.produce_arena("t", Arena.pack([b"x" * 1000, b"y"]))
    big = br.poll_arena(("t",), 10, 100)
    assert big.n == 1 and big.nbytes == 1000
