"""In-process broker with the surface calfkit calls on FastStream's KafkaBroker
(SURVEY.md §8b "downward" face): subscriber(*topics, group_id=, max_workers=, **kw),
publisher(topic, **kw)(handler), await publish(msg, topic=, correlation_id=, key=), start/stop and a
truthy `_connection` once started (reference calfkit/worker/worker.py:45-53, calfkit/nodes/base.py:82-87,
calfkit/client/base.py:137-147,164).

Kafka itself is out of scope (third-party transport; no client library or broker in
the image), so topics are in-memory queues of wire records; a real Kafka source/sink would implement
the same two methods the Worker uses: `poll_batch(topics, max_records)` and `produce_batch(...)`.
Records are bytes end to end: nothing here parses JSON.
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
    def __init__(self, *servers: Any, middlewares: list | None = None, **kwargs: Any):
        self.servers, self.kwargs = servers, kwargs
        self._connection: Any = None
        self.queues: dict[str, deque[Record]] = defaultdict(deque)
        self.subscriptions: list[Subscription] = []
        self.produced: int = 0

    # --- FastStream-shaped registration ---------------------------------------------------------
    def subscriber(self, *topics: str, group_id: str | None = None, max_workers: int = 1, **kwargs: Any) -> Subscription:
        sub = Subscription(tuple(topics), group_id, max_workers, kwargs)
        self.subscriptions.append(sub)
        return sub

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

    # --- consume side -------------------------------------------------------------------------------
    def poll_batch(self, topics: tuple[str, ...], max_records: int) -> list[Record]:
        out: list[Record] = []
        for t in topics:
            q = self.queues.get(t)
            while q and len(out) < max_records:
                out.append(q.popleft())
        return out

    def pending(self) -> int:
        return sum(len(q) for q in self.queues.values())


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
