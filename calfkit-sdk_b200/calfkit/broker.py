"""In-process broker with the surface calfkit calls on FastStream's KafkaBroker
(SURVEY.md §8b "downward" face): subscriber(*topics, group_id=, max_workers=, **kw),
publisher(topic, **kw)(handler), await publish(msg, topic=, correlation_id=, key=), start/stop and a
truthy `_connection` once started (reference calfkit/worker/worker.py:45-53, calfkit/nodes/base.py:82-87,
calfkit/client/base.py:137-147,164).

Kafka itself is out of scope (third-party transport; no client library or broker in
the image), so topics are in-memory queues of wire records; a real Kafka source/sink would implement
the same methods the Worker uses.  Records are bytes end to end: nothing here parses JSON.

Two granularities share the topics:
  * per-record (`Record`): the client's publishes, host-side nodes (LLM boundary, Python tools), tests;
  * per-batch: `produce_arena(topic, arena)` enqueues a whole page-locked batch (what a Kafka fetch lands, what a load
    generator pre-builds) and `produce_publishes(batch)` takes the payload arena + publish table a GPU lane produced and
    splits it per topic with vectorised index selections — no Python object per record.  `poll_arena` hands a consumer the
    next batch for its topics, bounded by records AND bytes (a batch never exceeds what an engine was sized for), zero-copy
    when the entry is contiguous.
Publishes to topics nobody subscribes to are counted and dropped when they arrive batch-wise (a long-running worker must
not grow without bound); per-record publishes are kept (tests read such topics afterwards).
Limitation (documented, not hidden): one shared queue per topic — consumer GROUPS are not modelled (two subscriptions to
one topic in different groups split the records instead of each seeing all of them, unlike Kafka).
"""
from __future__ import annotations

import asyncio
from collections import defaultdict, deque
from dataclasses import dataclass
from typing import Any, Awaitable, Callable


@dataclass
class Record:
    """One Kafka record as calfkit sees it: value bytes, optional key, the two headers FastStream sets."""
    topic: str
    value: bytes
    key: bytes | None = None
    correlation_id: str | None = None
    content_type: str = "application/json"


class Subscription:
    def __init__(self, topics: tuple[str, ...], group_id: str | None, max_workers: int, kwargs: dict[str, Any]):
        self.topics, self.group_id, self.max_workers, self.kwargs = topics, group_id, max_workers, kwargs
        self.handler: Callable[..., Awaitable[Any]] | None = None
        self.node = None                     # set by Worker for node subscriptions (batch path)
        self.publish_topic: str | None = None

    def __call__(self, handler):
        self.handler = handler
        self.node = getattr(handler, "__self__", None)
        return _HandlerRef(self, handler)


class _HandlerRef:
    """what `subscriber(handler)` returns; `publisher(topic)(ref)` attaches the return-value topic"""
    def __init__(self, sub: Subscription, handler):
        self.sub, self.handler = sub, handler

    def __call__(self, *a, **k):
        return self.handler(*a, **k)


class MemoryBroker:
    def __init__(self, *servers: Any, middlewares: list | None = None, num_partitions: int = 8, **kwargs: Any):
        self.servers, self.kwargs = servers, kwargs
        self.num_partitions = num_partitions       # partitions per topic (Kafka's default partitioner: murmur2(key) % partitions)
        self._connection: Any = None
        self.queues: dict[str, deque[Record]] = defaultdict(deque)
        self.arenas: dict[str, deque] = defaultdict(deque)            # topic -> deque[Arena]
        self.frags: dict[str, deque] = defaultdict(deque)             # topic -> deque[(PublishBatch, pub indices)]
        self.sinks: dict[str, Any] = {}                               # topic -> callable(PublishBatch, pub indices)
        self.subscriptions: list[Subscription] = []
        self.subscribed: set[str] = set()
        self.produced: int = 0
        self.dropped_unsubscribed: int = 0

    # --- FastStream-shaped registration ---------------------------------------------------------
    def subscriber(self, *topics: str, group_id: str | None = None, max_workers: int = 1, **kwargs: Any) -> Subscription:
        sub = Subscription(tuple(topics), group_id, max_workers, kwargs)
        self.subscriptions.append(sub)
        self.subscribed.update(topics)
        return sub

    def sink(self, topic: str, fn) -> None:
        """bulk consumer of one topic: fn(batch, pub_indices) is called from produce_publishes with the lane's batch (a
        Kafka producer would encode record batches from it; benchmarks count bytes).  The batch is valid during the call."""
        self.sinks[topic] = fn

    def publisher(self, topic: str, **kwargs: Any):
        def attach(ref: _HandlerRef):
            ref.sub.publish_topic = topic
            return ref
        return attach

    async def start(self) -> None:
        self._connection = object()

    async def stop(self) -> None:
        self._connection = None

    # --- produce side ------------------------------------------------------------------------------
    async def publish(self, msg: Any, topic: str, correlation_id: str | None = None, key: bytes | None = None, **kw: Any) -> None:
        if isinstance(msg, (bytes, bytearray, memoryview)):
            value = bytes(msg)
        elif hasattr(msg, "model_dump_json"):
            value = msg.model_dump_json().encode()       # object -> wire bytes at the user-API edge only
        else:
            raise TypeError(f"cannot publish {type(msg)!r}")
        self.produce(Record(topic, value, key, correlation_id))

    def produce(self, rec: Record) -> None:
        self.queues[rec.topic].append(rec)
        self.produced += 1

    def produce_batch(self, records: list[Record]) -> None:
        for r in records:
            self.produce(r)

    def produce_arena(self, topic: str, arena) -> None:
        """a whole batch of records for one topic (engine.lane.Arena)"""
        self.arenas[topic].append(arena)
        self.produced += arena.n

    def produce_publishes(self, batch) -> None:
        """split what a GPU lane produced (engine.lane.PublishBatch) over the topics, by index selection"""
        for tid, cnt in batch.topic_counts().items():
            self.produced += cnt
            if tid < 0 and batch.source is None:
                # forwarded by another rank without a registered topic id: the name stayed with the source record there.
                # Workers of one deployment register each other's topics (Worker(route_topics=...)): loud, not silent
                import logging
                logging.getLogger(__name__).error("%d forwarded publishes to a topic this worker has no id for: register it via route_topics", cnt)
                self.dropped_unsubscribed += cnt
                continue
            if tid < 0:
                # topics the engine has no id for (another worker's node, a client's private reply topic): the kernel left
                # the FNV-1a of the name in `pad`, so they are grouped vectorised and the name is decoded once per group
                import numpy as np
                idx = batch.select(-1)
                hashes = np.ascontiguousarray(batch.pubs["pad"][idx])
                for hv in np.unique(hashes):
                    sub = idx[hashes == hv]
                    tl = batch.pubs["topic_len"][sub]
                    if (tl != tl[0]).any():                     # hash collision between two names: per record
                        for topic, key, payload, _j in batch.iter_records(sub):
                            self.queues[topic].append(Record(topic, payload, key, key.decode() if key is not None else None))
                        continue
                    self._route(batch, batch.topic_name(int(sub[0])), sub)
                continue
            self._route(batch, batch.topic_names[tid], batch.select(tid))
        batch.release()

    def _route(self, batch, name: str, idx) -> None:
        if name in self.sinks:
            self.sinks[name](batch, idx)
        elif name in self.subscribed:
            self.frags[name].append((batch.retain(), idx))
        else:
            self.dropped_unsubscribed += len(idx)

    # --- consume side -------------------------------------------------------------------------------
    def poll_batch(self, topics: tuple[str, ...], max_records: int, max_bytes: int | None = None) -> list[Record]:
        """per-record consume, bounded by records and (optionally) bytes; batch-wise entries are materialised as Records"""
        out: list[Record] = []
        nbytes = 0
        for t in topics:
            q = self.queues.get(t)
            while q and len(out) < max_records:
                if max_bytes is not None and out and nbytes + len(q[0].value) > max_bytes:
                    return out
                nbytes += len(q[0].value)
                out.append(q.popleft())
            if len(out) < max_records and (self.arenas.get(t) or self.frags.get(t)):
                arena = self.poll_arena((t,), max_records - len(out), None if max_bytes is None else max(max_bytes - nbytes, 1),
                                        records_too=False)
                if arena is not None:
                    for i in range(arena.n):
                        out.append(Record(t, arena.record(i)))
                    nbytes += arena.nbytes
                    arena.release()
        return out

    def poll_arena(self, topics: tuple[str, ...], max_records: int, max_bytes: int | None = None, records_too: bool = True):
        """the next batch for these topics as one contiguous Arena (None when idle).  Never returns more than max_records
        records or max_bytes bytes (one oversized record is still returned alone: the engine reports it per record)."""
        from calfkit.engine.lane import Arena
        for t in topics:
            q = self.arenas.get(t)
            if q:
                a = q[0]
                k = _fit(a.offsets, max_records, max_bytes)
                if k >= a.n:
                    return q.popleft()
                head = a.slice(0, k)
                rest = a.slice(k, a.n)
                rest._on_release, a._on_release = a._on_release, None       # the tail keeps the buffer alive
                q[0] = rest
                return head
            f = self.frags.get(t)
            if f:
                batch, idx = f[0]
                lens = batch.out_len[batch.pubs["payload"][idx]].astype("int64")
                import numpy as np
                k = _fit(np.concatenate(([0], np.cumsum(lens))), max_records, max_bytes)
                arena = batch.gather(idx[:k])        # a zero-copy view retains the batch until the arena is released
                if k >= len(idx):
                    f.popleft()
                    batch.release()
                else:
                    f[0] = (batch, idx[k:])
                return arena
            if records_too and self.queues.get(t):
                recs = self.poll_batch((t,), max_records, max_bytes)
                return Arena.pack([r.value for r in recs])
        return None

    def pending(self) -> int:
        return (sum(len(q) for q in self.queues.values()) + sum(a.n for q in self.arenas.values() for a in q)
                + sum(len(i) for q in self.frags.values() for _b, i in q))


def _fit(offsets, max_records: int, max_bytes: int | None) -> int:
    """how many leading records of a batch fit the limits (at least one)"""
    n = len(offsets) - 1
    k = min(n, max_records)
    if n and (max_bytes is None or int(offsets[k] - offsets[0]) <= max_bytes):
        return k                                            # the common case: no array work at all
    if max_bytes is not None and n:
        import numpy as np
        k = min(k, int(np.searchsorted(offsets - offsets[0], max_bytes, side="right")) - 1)
    return max(k, 1) if n else 0


# the names calfkit code imports from faststream.kafka
KafkaBroker = MemoryBroker


class TestKafkaBroker:
    """async context manager mirroring faststream's TestKafkaBroker: starts the in-memory broker"""
    def __init__(self, broker: MemoryBroker):
        self.broker = broker

    async def __aenter__(self) -> MemoryBroker:
        await self.broker.start()
        return self.broker

    async def __aexit__(self, *exc: object) -> None:
        await self.broker.stop()


async def _maybe_await(x):
    if asyncio.iscoroutine(x):
        return await x
    return x
