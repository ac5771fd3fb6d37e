"""oracle/kafka_batch.py (RecordBatch v2 framing + CRC32C) against published known-answer vectors.
The Kafka client library the reference uses (aiokafka, via FastStream) is absent from the image: the pins are the RFC 3720
CRC32C vectors and a complete v2 frame from kafka-python's test suite whose stored CRC must match our computation."""
import struct

from oracle import kafka_batch as kb

# kafka-python test/record/test_default_records.py: record_batch_data_v2[0] — one record, no key, value b"123"
KAFKA_PYTHON_V2_FRAME = (
    b'\x00\x00\x00\x00\x00\x00\x00\x00\x00\x00\x00;\x00\x00\x00\x01\x02\x03'
    b'\x18\xa2p\x00\x00\x00\x00\x00\x00\x00\x00\x01]\xff{\x06<\x00\x00\x01]'
    b'\xff{\x06<\xff\xff\xff\xff\xff\xff\xff\xff\xff\xff\xff\xff\xff\xff'
    b'\x00\x00\x00\x01\x12\x00\x00\x00\x01\x06123\x00')


def test_crc32c_rfc3720_vectors():
    assert kb.crc32c(b"123456789") == 0xE3069283
    assert kb.crc32c(bytes(32)) == 0x8A9136AA
    assert kb.crc32c(b"\xff" * 32) == 0x62A8AB43
    assert kb.crc32c(bytes(range(32))) == 0x46DD794E
    assert kb.crc32c(bytes(range(31, -1, -1))) == 0x113FDB5C
    # incremental form
    assert kb.crc32c(b"6789", kb.crc32c(b"12345")) == 0xE3069283


def test_published_v2_frame_crc_and_decode():
    f = KAFKA_PYTHON_V2_FRAME
    assert len(f) == 12 + struct.unpack_from(">i", f, 8)[0]
    stored = struct.unpack_from(">I", f, 17)[0]
    assert stored == 0x0318A270 and kb.crc32c(f[kb.CRC_FROM:]) == stored
    recs = kb.decode_batches(f)
    assert len(recs) == 1 and recs[0].value == b"123" and recs[0].key is None and recs[0].headers == []
    assert recs[0].offset == 0 and recs[0].timestamp == 1503229838908
    # re-encoding the decoded record with the frame's own header fields reproduces the frame byte for byte
    again = kb.encode_batch([kb.KRecord(b"123", None, [], 0, 0)], base_offset=0, base_timestamp=1503229838908, partition_leader_epoch=1)
    assert again == f


def test_varints():
    for n, enc in ((0, b"\x00"), (-1, b"\x01"), (1, b"\x02"), (63, b"\x7e"), (64, b"\x80\x01"), (-65, b"\x81\x01"), (300, b"\xd8\x04"),
                   (2147483647, b"\xfe\xff\xff\xff\x0f"), (-2147483648, b"\xff\xff\xff\xff\x0f")):
        assert kb.put_varint(n) == enc and kb.get_varint(enc, 0) == (n, len(enc))


def test_round_trip_calfkit_records():
    recs = [kb.calfkit_record(b'{"a":%d}' % i, f"{i:032x}", keyed=bool(i % 3)) for i in range(200)] + [kb.KRecord(None, b"k"), kb.KRecord(b"", None)]
    buf = kb.encode_batch(recs[:120], base_offset=1000, base_timestamp=1700000000000) + kb.encode_batch(recs[120:], base_offset=1120)
    out = kb.decode_batches(buf)
    assert [(r.value, r.key, r.headers) for r in out] == [(r.value, r.key, r.headers) for r in recs]
    assert [r.offset for r in out] == list(range(1000, 1202))
    # a flipped payload bit is caught by the CRC
    bad = bytearray(buf); bad[200] ^= 1
    try:
        kb.decode_batches(bytes(bad))
        raise AssertionError("corruption not detected")
    except ValueError:
        pass
    # a truncated trailing batch (legal in a fetch response) is ignored
    assert len(kb.decode_batches(buf[:-5])) == 120
