"""Worker (mirrors reference calfkit/worker/worker.py:12-61).

Same constructor, add_nodes, register_handlers (idempotence guard -> RuntimeError) and async run().
The reference's run() hands control to FastStream, which then delivers ONE record at a time to
node.handler.  Here run() is the batch loop itself: poll up to `batch_records` records per node from
the broker, push them through the node's CUDA plan (node.process_batch -> BatchEngine), produce the
routed outputs, repeat.  One BatchEngine (one CUDA stream, one set of HBM buffers) per node.
"""
from __future__ import annotations

import asyncio
import logging
from typing import Any

from calfkit.broker import Record
from calfkit.client import Client
from calfkit.engine import BatchEngine
from calfkit.nodes import BaseNodeDef

logger = logging.getLogger(__name__)


def engine_for(node: BaseNodeDef, *, device: int = 0, max_records: int = 1 << 14, max_in_bytes: int = 64 << 20,
               extra_topics: list[str] | None = None) -> BatchEngine:
    """lazily created, per-node engine (also used by the object-level BaseNodeDef.handler)"""
    eng = getattr(node, "_ck_engine", None)     # the engine lives (and dies) with its node: no id()-keyed registry
    if eng is None:
        eng = BatchEngine(device, max_records=max_records, max_in_bytes=max_in_bytes,
                          max_out_bytes=8 * max_in_bytes, max_payloads=8 * max_records)
        topics = list(node.subscribe_topics) + ([node.publish_topic] if node.publish_topic else [])
        for t in getattr(node, "tools", []) or []:
            topics += list(t.subscribe_topics)
        eng.register_topics(topics + (extra_topics or []), num_partitions=0)
        node.configure_engine(eng)
        node._ck_engine = eng
    return eng


class Worker:
    def __init__(self, client: Client, nodes: list[BaseNodeDef] | None = None, max_workers: int = 1,
                 group_id: str | None = None, extra_publish_kwargs: dict[str, Any] | None = None,
                 extra_subscribe_kwargs: dict[str, Any] | None = None, *, device: int = 0, batch_records: int = 1 << 14):
        self._client = client
        self._nodes = nodes or list()
        self._max_workers = max_workers
        self._group_id = group_id
        self._extra_publish_kwargs = extra_publish_kwargs or {}
        self._extra_subscribe_kwargs = extra_subscribe_kwargs or {}
        self._prepared = False
        self._device = device
        self._batch_records = batch_records
        self._subs: list[tuple[BaseNodeDef, Any]] = []

    def add_nodes(self, *nodes: BaseNodeDef) -> None:
        self._nodes.extend(nodes)

    def register_handlers(self) -> None:
        if self._prepared:
            raise RuntimeError("register_handlers() already called")
        for node in self._nodes:
            group_id = self._group_id or node.name
            logger.info("registering node=%s subscribe=%s publish=%s", node.name, node.subscribe_topics, node.publish_topic)
            subscriber = self._client._connection.subscriber(*node.subscribe_topics, group_id=group_id,
                                                             max_workers=self._max_workers, **self._extra_subscribe_kwargs)
            handler = subscriber(node.handler)
            if node.publish_topic:
                self._client._connection.publisher(node.publish_topic, **self._extra_publish_kwargs)(handler)
            self._subs.append((node, subscriber))
        self._prepared = True

    def step(self) -> int:
        """one pass over all nodes: poll -> CUDA batch -> produce.  Returns the number of records consumed."""
        broker = self._client._connection
        consumed = 0
        for node, sub in self._subs:
            records: list[Record] = broker.poll_batch(sub.topics, self._batch_records)
            if not records:
                continue
            consumed += len(records)
            engine = engine_for(node, device=self._device, max_records=self._batch_records)
            broker.produce_batch(node.process_batch(engine, records))
        return consumed

    async def run(self, *, until_idle: bool = False, idle_sleep: float = 0.001, **extra_run_args: Any) -> None:
        """Run the worker as a service (reference: blocks in FastStream(...).run()); `until_idle=True`
        returns once every subscribed topic is drained — used by tests and the config-1 example."""
        logger.info("worker starting with %d node(s)", len(self._nodes))
        if not self._prepared:
            self.register_handlers()
        broker = self._client._connection
        if not broker._connection:
            await broker.start()
        while True:
            n = self.step()
            n += await self._client._dispatcher.drain(broker) if hasattr(self._client, "_dispatcher") else 0
            if n == 0:
                if until_idle:
                    return
                await asyncio.sleep(idle_sleep)
            else:
                await asyncio.sleep(0)
