"""SURVEY.md section 8f row 1 on the device: a raw fetch response (concatenated Kafka RecordBatch v2 frames built by the oracle
restatement oracle/kafka_batch.py, itself pinned to published known-answer vectors in tests/test_kafka_batch.py) feeds the
walker directly; the produce side builds frames on the device that must equal the oracle's byte for byte."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture()
def engine():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from calfkit.engine import BatchEngine, ToolTemplate
    e = BatchEngine(0, max_records=1 << 15, max_in_bytes=64 << 20)
    e.register_topics(["tool.get_weather.input", "tool.get_weather.output", "weather_agent.input"], num_partitions=8)
    e.set_tool_node("tool.get_weather.output", ToolTemplate.from_format("It's sunny in {location}"))
    yield e
    e.close()


def frames_of(recs, sizes, corrupt=None, truncate_tail=False):
    from oracle import kafka_batch as kb
    import json
    out, k, krecs = [], 0, []
    for bi, sz in enumerate(sizes):
        chunk = []
        for r in recs[k:k + sz]:
            corr = json.loads(r)["context"]["deps"]["correlation_id"]
            chunk.append(kb.calfkit_record(r, corr, keyed=(len(chunk) % 4 != 0)))     # every 4th: the client's unkeyed first publish
        krecs += chunk
        f = bytearray(kb.encode_batch(chunk, base_offset=1000 + k, base_timestamp=1767225600000 + bi))
        if corrupt == bi:
            f[len(f) // 2] ^= 0x40
        out.append(bytes(f))
        k += sz
    buf = b"".join(out)
    if truncate_tail:
        buf += out[0][:37]                                      # a fetch response may end in the middle of a frame
    return buf, krecs


def test_crc32c_known_answers_on_device(engine):
    """frames whose only record value is a known-answer input: the device must accept exactly the frames whose stored CRC
    is right (the oracle computes it, and is itself pinned to RFC 3720 / kafka-python vectors)"""
    from oracle import kafka_batch as kb
    from test_kafka_batch import KAFKA_PYTHON_V2_FRAME
    good = [KAFKA_PYTHON_V2_FRAME, kb.encode_batch([kb.KRecord(b"123456789")]), kb.encode_batch([kb.KRecord(bytes(range(256)) * 300)]),
            kb.encode_batch([kb.KRecord(b"x" * n) for n in (0, 1, 2, 3, 4, 5, 63, 64, 65, 127, 128, 129, 4095, 4096, 4097)])]
    for f in good:
        for flip in (None, 21, len(f) - 1, len(f) // 2, 18):
            g = bytearray(f)
            if flip is not None:
                g[flip] ^= 1
            engine.submit_recordbatch(np.frombuffer(bytes(g), dtype=np.uint8))
            bad = engine.rb_index()["bad"]
            assert len(bad) == len(kb.decode_batches(f)) and (bad == (0 if flip is None else 1)).all(), (len(f), flip)


def test_fetch_response_feeds_the_walker(engine):
    import tools_def
    from calfkit import synth
    from calfkit.engine._lib import CK_BAD_FRAME, CK_OK, COL
    from oracle import kafka_batch as kb, port
    recs = synth.tool_events(1500, seed=61) + synth.tool_events(200, seed=62, size=None, full_history=True) + synth.mixed_events(100, seed=63, hi=30000)
    sizes = [1, 14, 14, 200, 3, 500, 700, 268, 100]
    assert sum(sizes) == len(recs)
    buf, krecs = frames_of(recs, sizes, corrupt=4, truncate_tail=True)
    lo, hi = sum(sizes[:4]), sum(sizes[:5])                     # the records of the corrupted frame (one flipped bit)
    dec = [r.value for r in kb.decode_batches(buf[:len(buf) - 37], verify_crc=False)]
    assert dec[:lo] == recs[:lo] and dec[hi:] == recs[hi:]
    n = engine.submit_recordbatch(np.frombuffer(buf, dtype=np.uint8))
    assert n == len(recs)
    ix = engine.rb_index()
    assert (ix["bad"][lo:hi] == 1).all() and ix["bad"].sum() == hi - lo
    for i in list(range(0, lo, 37)) + list(range(hi, n, 53)):
        k = krecs[i]
        assert buf[ix["val_off"][i]:ix["val_off"][i] + ix["val_len"][i]] == k.value
        assert (ix["key_len"][i] == -1) == (k.key is None)
        if k.key is not None:
            assert buf[ix["key_off"][i]:ix["key_off"][i] + ix["key_len"][i]] == k.key
        assert buf[ix["corr_off"][i]:ix["corr_off"][i] + ix["corr_len"][i]] == dict(k.headers)["correlation_id"]
    engine.tool_plan()
    out = engine.fetch()
    st = out.cols[COL["STATUS"]]
    assert (st[lo:hi] == CK_BAD_FRAME).all() and (np.delete(st, np.s_[lo:hi]) == CK_OK).all()
    node = port.ToolNode.of(tools_def.get_weather)
    got = [(p.topic, p.key, p.payload) for p in out.publishes()]
    want = [(t, k, pl) for i, r in enumerate(recs) if not lo <= i < hi for (t, k, _c, pl) in port.tool_node_event(node, r)]
    assert got == want


def test_produce_frames_match_oracle(engine):
    from calfkit import synth
    from calfkit.engine._lib import COL
    from oracle import kafka_batch as kb
    recs = synth.tool_events(4000, seed=71) + synth.tool_events(300, seed=72, size=None, full_history=True)
    b = synth.pack(recs)
    engine.submit(b.data, b.offsets)
    engine.tool_plan()
    out = engine.fetch()
    tid = engine.topic_ids["weather_agent.input"]
    for part in (0, 3, 7):
        idx = np.nonzero((out.pubs["payload"] != 0xFFFFFFFF) & (out.pubs["topic_id"] == tid) & (out.pubs["partition"] == part))[0]
        assert len(idx) > 100
        frame = engine.encode_recordbatch(idx, base_offset=5000 + part, timestamp_ms=1767225600123).tobytes()
        want = kb.encode_batch([kb.calfkit_record(out.payload(int(out.pubs["payload"][j])), out.key_of(out.pubs[j]).decode()) for j in idx],
                               base_offset=5000 + part, base_timestamp=1767225600123)
        assert frame == want
    # the unkeyed handler-return publishes (worker/worker.py:52-53): null key, still both headers
    tid2 = engine.topic_ids["tool.get_weather.output"]
    idx = np.nonzero((out.pubs["payload"] != 0xFFFFFFFF) & (out.pubs["topic_id"] == tid2))[0][:777]
    frame = engine.encode_recordbatch(idx, base_offset=9, timestamp_ms=1).tobytes()
    corr = lambda j: out.record_bytes(int(out.pubs["record"][j]))[int(out.cols[COL["CORR_OFF"], out.pubs["record"][j]]):][:int(out.cols[COL["CORR_LEN"], out.pubs["record"][j]])].tobytes().decode()  # noqa: E731
    want = kb.encode_batch([kb.calfkit_record(out.payload(int(out.pubs["payload"][j])), corr(j), keyed=False) for j in idx], base_offset=9, base_timestamp=1)
    assert frame == want
    # and the device reads its own frames back
    n = engine.submit_recordbatch(np.frombuffer(frame, dtype=np.uint8))
    assert n == len(idx) and engine.rb_index()["bad"].sum() == 0
