"""Worker's host-side node path without a GPU: a fake node stands in for Python tools / the Agent's LLM boundary (their
process_batch needs an engine; here the engine is a stub).  What is checked is the Worker's own contract:
polls are bounded by records AND bytes, a record that was dequeued is never lost to a batch-level failure (the batch is
split and retried), and a record that fails alone is dropped alone."""
import asyncio

import pytest

from calfkit import Client, Worker
from calfkit.broker import Record
from calfkit.nodes import BaseNodeDef


class FakeNode(BaseNodeDef):
    """consumes `fake.in`, publishes one record per input to `fake.out`; refuses batches above `max_batch` records (an engine
    whose buffers are too small) and any batch that contains a poisoned record"""
    def __init__(self, max_batch: int):
        self.node_id = "fake"
        self.subscribe_topics = ["fake.in"]
        self.publish_topic = "fake.out"
        self.max_batch = max_batch
        self.calls: list[int] = []

    async def run(self, *a, **k):       # the object-level surface is not used here
        raise NotImplementedError

    def process_batch(self, engine, records, **kw):
        self.calls.append(len(records))
        if len(records) > self.max_batch:
            raise RuntimeError("batch does not fit the engine")
        if any(r.value == b"poison" for r in records):
            raise ValueError("user code raised outside the per-record guard")
        return [Record("fake.out", r.value.upper(), r.key, r.correlation_id) for r in records]


@pytest.fixture()
def setup(monkeypatch):
    import calfkit.worker.worker as W
    monkeypatch.setattr(W, "engine_for", lambda node, **kw: object())
    client = Client.connect()
    return client, client._connection


def _run(worker):
    asyncio.run(worker.run(until_idle=True))


def test_batches_are_split_and_no_dequeued_record_is_lost(setup):
    client, broker = setup
    node = FakeNode(max_batch=4)
    worker = Worker(client, nodes=[node], batch_records=16, batch_bytes=1 << 20)
    vals = [f"r{i:02d}".encode() for i in range(37)]
    for v in vals:
        broker.produce(Record("fake.in", v))
    _run(worker)
    out = sorted(r.value for r in broker.queues["fake.out"])
    assert out == sorted(v.upper() for v in vals)                     # every record came out exactly once
    assert max(node.calls) == 16 and 4 in node.calls                  # polled 16 at a time, halved down to what fits
    assert worker.stats["records"] == 37


def test_a_record_that_fails_alone_is_dropped_alone(setup):
    client, broker = setup
    node = FakeNode(max_batch=64)
    worker = Worker(client, nodes=[node], batch_records=64, batch_bytes=1 << 20)
    vals = [b"a", b"b", b"poison", b"c", b"d", b"e", b"poison", b"f"]
    for v in vals:
        broker.produce(Record("fake.in", v))
    _run(worker)
    assert sorted(r.value for r in broker.queues["fake.out"]) == [b"A", b"B", b"C", b"D", b"E", b"F"]
    assert 1 in node.calls                                            # narrowed down to the single failing records


def test_polls_honour_the_byte_bound(setup):
    client, broker = setup
    node = FakeNode(max_batch=1000)
    worker = Worker(client, nodes=[node], batch_records=1000, batch_bytes=250)
    for i in range(10):
        broker.produce(Record("fake.in", bytes([65 + i]) * 100))
    _run(worker)
    assert len(broker.queues["fake.out"]) == 10
    assert max(node.calls) == 2                                       # two 100-byte records per poll: never more than 250 bytes
