"""TEST INFRASTRUCTURE ONLY — CPU restatement of Kafka's RecordBatch v2 framing (magic 2) and CRC32C.

Where this sits in the reference: calfkit never touches the framing itself; it is done by aiokafka underneath
`broker.subscriber(...)` / `broker.publisher(...)` / `broker.publish(...)` (reference calfkit/worker/worker.py:45-53,
calfkit/nodes/base.py:82-87), pulled in transitively by `faststream[kafka]>=0.6.6` (pyproject.toml:23, unpinned; no
lockfile).  THIRD-PARTY SOURCE ABSENT: neither aiokafka nor kafka-python nor confluent-kafka is in /root/reference or in
the image, so this file restates the *published* format (Apache Kafka protocol guide, "Record Batch" / "Record", message
format v2, KIP-98) and is pinned to published known-answer vectors instead of to the library:

  * CRC32C (Castagnoli, reflected polynomial 0x82F63B78): the RFC 3720 appendix B.4 vectors (32 x 00, 32 x FF,
    ascending, descending) and the classic check value of "123456789" = 0xE3069283;
  * a complete RecordBatch v2 frame (one record, value "123") from kafka-python's test suite
    (test/record/test_default_records.py `record_batch_data_v2[0]`): its stored CRC 0x0318A270 must equal the CRC32C this
    file computes over the covered bytes, and decoding must give back the record — see tests/test_kafka_batch.py.

Layout (all fixed-width integers big-endian):
  baseOffset i64 | batchLength i32 | partitionLeaderEpoch i32 | magic i8 (=2) | crc u32 |            <- 21 bytes, not covered
  attributes i16 | lastOffsetDelta i32 | baseTimestamp i64 | maxTimestamp i64 | producerId i64 |
  producerEpoch i16 | baseSequence i32 | recordsCount i32 | records...                               <- covered by crc
  record:  length varint | attributes i8 | timestampDelta varlong | offsetDelta varint | keyLength varint | key |
           valueLength varint | value | headersCount varint | { keyLength varint | key | valueLength varint | value }*
  varints are zig-zag encoded, 7 bits per byte, little-endian groups; length -1 means null.
What calfkit puts in a record: value = Envelope JSON, key = correlation_id bytes (absent on the
client's first publish), headers `correlation_id` and `content-type: application/json` (FastStream).
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field

HEADER_LEN = 61          # bytes before the first record
CRC_FROM = 21            # the crc covers [21, end of batch)

_POLY = 0x82F63B78
_TABLE = []
for _i in range(256):
    _c = _i
    for _ in range(8):
        _c = (_c >> 1) ^ (_POLY if _c & 1 else 0)
    _TABLE.append(_c)


def crc32c(data: bytes, crc: int = 0) -> int:
    c = crc ^ 0xFFFFFFFF
    for b in data:
        c = _TABLE[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def zigzag(n: int) -> int:
    return (n << 1) ^ (n >> 63)


def unzigzag(u: int) -> int:
    return (u >> 1) ^ -(u & 1)


def put_varint(n: int) -> bytes:
    u = zigzag(n) & 0xFFFFFFFFFFFFFFFF
    out = bytearray()
    while u >= 0x80:
        out.append((u & 0x7F) | 0x80)
        u >>= 7
    out.append(u)
    return bytes(out)


def get_varint(buf: bytes, pos: int) -> tuple[int, int]:
    u = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        u |= (b & 0x7F) << shift
        if not b & 0x80:
            return unzigzag(u), pos
        shift += 7
        if shift > 63:
            raise ValueError("varint too long")


@dataclass
class KRecord:
    value: bytes | None
    key: bytes | None = None
    headers: list[tuple[str, bytes | None]] = field(default_factory=list)
    offset: int = 0
    timestamp: int = 0


def encode_record(r: KRecord, offset_delta: int, ts_delta: int) -> bytes:
    body = bytearray(b"\x00")                                   # attributes
    body += put_varint(ts_delta) + put_varint(offset_delta)
    body += put_varint(-1) if r.key is None else put_varint(len(r.key)) + r.key
    body += put_varint(-1) if r.value is None else put_varint(len(r.value)) + r.value
    body += put_varint(len(r.headers))
    for k, v in r.headers:
        kb = k.encode()
        body += put_varint(len(kb)) + kb
        body += put_varint(-1) if v is None else put_varint(len(v)) + v
    return put_varint(len(body)) + bytes(body)


def encode_batch(records: list[KRecord], base_offset: int = 0, base_timestamp: int = 0, partition_leader_epoch: int = 0,
                 producer_id: int = -1, producer_epoch: int = -1, base_sequence: int = -1) -> bytes:
    """one uncompressed RecordBatch v2 frame; record i gets offsetDelta i and timestampDelta (timestamp - base)"""
    recs = b"".join(encode_record(r, i, (r.timestamp - base_timestamp) if r.timestamp else 0) for i, r in enumerate(records))
    max_ts = max([r.timestamp for r in records if r.timestamp] or [base_timestamp])
    covered = struct.pack(">hiqqqhii", 0, len(records) - 1, base_timestamp, max_ts, producer_id, producer_epoch, base_sequence,
                          len(records)) + recs
    head = struct.pack(">qiibI", base_offset, len(covered) + 9, partition_leader_epoch, 2, crc32c(covered))
    return head + covered


def decode_batches(buf: bytes, verify_crc: bool = True) -> list[KRecord]:
    """every record of a concatenation of RecordBatch v2 frames (a fetch response's record set), in order"""
    out: list[KRecord] = []
    pos = 0
    while pos + 12 <= len(buf):
        base_offset, batch_len = struct.unpack_from(">qi", buf, pos)
        end = pos + 12 + batch_len
        if end > len(buf):
            break                                               # a truncated trailing batch is legal in a fetch response
        _epoch, magic, crc = struct.unpack_from(">ibI", buf, pos + 12)
        if magic != 2:
            raise ValueError(f"magic {magic} is not a v2 batch")
        if verify_crc and crc32c(buf[pos + CRC_FROM:end]) != crc:
            raise ValueError("CRC32C mismatch")
        attrs, _last, base_ts, _max_ts, _pid, _pep, _seq, count = struct.unpack_from(">hiqqqhii", buf, pos + CRC_FROM)
        if attrs & 7:
            raise ValueError("compressed batches are not handled")
        p = pos + HEADER_LEN
        for _ in range(count):
            length, p = get_varint(buf, p)
            rec_end = p + length
            p += 1                                              # attributes
            ts_delta, p = get_varint(buf, p)
            off_delta, p = get_varint(buf, p)
            klen, p = get_varint(buf, p)
            key = None
            if klen >= 0:
                key, p = buf[p:p + klen], p + klen
            vlen, p = get_varint(buf, p)
            value = None
            if vlen >= 0:
                value, p = buf[p:p + vlen], p + vlen
            nh, p = get_varint(buf, p)
            headers = []
            for _h in range(nh):
                hk, p = get_varint(buf, p)
                hkey, p = buf[p:p + hk].decode(), p + hk
                hv, p = get_varint(buf, p)
                hval = None
                if hv >= 0:
                    hval, p = buf[p:p + hv], p + hv
                headers.append((hkey, hval))
            if p != rec_end:
                raise ValueError("record length mismatch")
            out.append(KRecord(value, key, headers, base_offset + off_delta, base_ts + ts_delta))
        pos = end
    return out


def calfkit_record(value: bytes, correlation_id: str | None, keyed: bool = True) -> KRecord:
    """a record as calfkit's publishes look on the wire (nodes/base.py:82-87; FastStream adds the two headers)"""
    headers: list[tuple[str, bytes | None]] = [("content-type", b"application/json")]
    if correlation_id is not None:
        headers.append(("correlation_id", correlation_id.encode()))
    return KRecord(value, correlation_id.encode() if (keyed and correlation_id is not None) else None, headers)
