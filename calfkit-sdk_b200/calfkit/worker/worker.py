"""Worker (mirrors reference calfkit/worker/worker.py:12-61).

Same constructor, add_nodes, register_handlers (idempotence guard -> RuntimeError) and async run().
The reference's run() hands control to FastStream, which then delivers ONE record at a time to
node.handler.  Here run() is the batch loop itself.  Two paths per node:

  * device-template tool nodes (the data-parallel hot path): `broker.poll_arena` -> one pinned batch arena ->
    `LanePipeline.push` (H2D + decode + plan + encode + route on one of K lanes while an older step's results come
    back) -> `broker.produce_publishes(PublishBatch)`.  No Python object per record.
  * nodes with a host half (Python tools, the Agent's LLM boundary): `broker.poll_batch` -> `node.process_batch`
    (engine for decode / plan / encode, Python only for the user callable) -> `broker.produce_batch`.

A poll is bounded by records AND bytes, an engine-level failure splits the batch and retries (dequeued records are never
lost), and records the device declines as CK_UNSUPPORTED are re-run through the host-tool path.
"""
from __future__ import annotations

import asyncio
import logging
import os
from typing import Any

import numpy as np

from calfkit.broker import Record
from calfkit.client import Client
from calfkit.engine import BatchEngine
from calfkit.engine._lib import CK_ACT_RAISES, CK_OK, CK_UNSUPPORTED, STATUS_NAMES
from calfkit.engine.lane import LanePipeline, PublishBatch
from calfkit.nodes import BaseNodeDef

logger = logging.getLogger(__name__)


def _EMPTY_ARENA():
    from calfkit.engine.lane import Arena
    return Arena(np.zeros(0, dtype=np.uint8), np.zeros(1, dtype=np.int64))


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
                 extra_subscribe_kwargs: dict[str, Any] | None = None, *, device: int | None = None,
                 batch_records: int = 1 << 14, batch_bytes: int = 64 << 20, lanes: int = 3,
                 route_topics: list[str] | None = None, rank: int | None = None, world: int | None = None):
        """Reference signature (worker/worker.py:13-31) plus keyword-only engine knobs: `device` (default: LOCAL_RANK of a
        one-process-per-GPU launch, else 0), the batch bounds a poll honours (records AND bytes: a batch never exceeds what
        the engines were sized for), the number of pipelined lanes per device-template node, and `route_topics`: names of
        topics owned by OTHER workers that this worker's outputs go to (agents' input topics ...) so that the device resolves
        them to ids too; names it does not know are still routed, grouped by hash on the host.
        `rank` / `world` (default: torch.distributed's, when initialised): one worker process per GPU, records sharded by Kafka
        partition (reference: consumer-group sharding, worker/worker.py:38,47, with every hop keyed by correlation_id,
        nodes/base.py:86,103,117,134).  A keyed publish whose partition (murmur2(key) % partitions) is owned by another rank
        (partition % world) is stored straight into that rank's receive buffer over NVLink and produced THERE; all ranks
        tick in lockstep (an idle rank submits an empty batch) so that the per-step barriers of the exchange match."""
        self._client = client
        self._nodes = nodes or list()
        self._max_workers = max_workers
        self._group_id = group_id
        self._extra_publish_kwargs = extra_publish_kwargs or {}
        self._extra_subscribe_kwargs = extra_subscribe_kwargs or {}
        self._prepared = False
        self._device = device if device is not None else int(os.environ.get("LOCAL_RANK", 0))
        self._batch_records = batch_records
        self._batch_bytes = batch_bytes
        self._lanes = lanes
        self._route_topics = list(route_topics or [])
        if rank is None or world is None:
            try:
                import torch.distributed as dist
                if dist.is_available() and dist.is_initialized():
                    rank, world = dist.get_rank(), dist.get_world_size()
            except ImportError:
                pass
        self._rank, self._world = (rank or 0), (world or 1)
        self._subs: list[tuple[BaseNodeDef, Any]] = []
        self._pipes: dict[int, LanePipeline] = {}
        self.stats = {"records": 0, "publishes": 0, "rejected": 0, "raises": 0, "host_fallback": 0, "steps": 0}

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

    # ---- device-template tool nodes: arena in -> pipelined lanes -> publish batch out, no Python per record ----------
    def _pipeline(self, node: BaseNodeDef) -> LanePipeline:
        pipe = self._pipes.get(id(node))
        if pipe is None:
            topics = list(node.subscribe_topics) + ([node.publish_topic] if node.publish_topic else []) + self._route_topics

            def configure(eng: BatchEngine) -> None:
                eng.register_topics(topics, num_partitions=getattr(self._client._connection, "num_partitions", 0))
                node.configure_engine(eng)
            pipe = LanePipeline(self._device, configure, lanes=self._lanes, max_records=self._batch_records,
                                max_in_bytes=self._batch_bytes, exchange=(self._rank, self._world) if self._world > 1 else None)
            self._pipes[id(node)] = pipe
        return pipe

    def _produce_fast(self, node: BaseNodeDef, batch: PublishBatch) -> None:
        broker = self._client._connection
        n = batch.source.n
        if n == 0:
            batch.release()
            return
        self.stats["records"] += n
        bad = batch.status != CK_OK
        nbad = int(np.count_nonzero(bad))
        if nbad:
            self.stats["rejected"] += nbad
            for i in np.nonzero(bad & (batch.status != CK_UNSUPPORTED))[0][:8]:
                logger.error("record %d rejected: %s", int(i), STATUS_NAMES[int(batch.status[i])])
            # the device declined these (a construct the template / fast path does not cover, e.g. `args` given as a JSON
            # string or a non-string template argument): the host-tool path handles them, nothing is dropped silently
            redo = np.nonzero((batch.status == CK_UNSUPPORTED) & (batch.action == CK_ACT_RAISES))[0]
            if len(redo):
                self.stats["host_fallback"] += len(redo)
                recs = [Record(node.subscribe_topics[0], batch.source.record(int(i))) for i in redo]
                self._run_host(node, recs, force_host=True)
        self.stats["raises"] += int(np.count_nonzero((batch.action == CK_ACT_RAISES) & ~bad))
        self.stats["publishes"] += batch.n_publishes
        broker.produce_publishes(batch)

    # ---- host-side nodes (Python tools, the Agent's LLM boundary): per-record objects are inherent there -----------
    def _run_host(self, node: BaseNodeDef, records: list[Record], **kw: Any) -> None:
        """never loses a dequeued record: an engine-level failure (a batch the buffers cannot hold) splits the batch and
        retries; a single record that still fails is logged and dropped alone"""
        broker = self._client._connection
        engine = engine_for(node, device=self._device, max_records=self._batch_records, max_in_bytes=self._batch_bytes)
        try:
            broker.produce_batch(node.process_batch(engine, records, **kw) if kw else node.process_batch(engine, records))
            self.stats["records"] += 0 if kw else len(records)
        except Exception:  # noqa: BLE001  (EngineError: capacity; user code outside the per-record guards)
            if len(records) == 1:
                logger.exception("node %s: record dropped after failing alone", node.name)
                return
            mid = len(records) // 2
            logger.warning("node %s: batch of %d failed, retrying in halves", node.name, len(records))
            self._run_host(node, records[:mid], **kw)
            self._run_host(node, records[mid:], **kw)

    def step(self) -> int:
        """one pass over all nodes: poll -> CUDA batch -> produce.  Returns the number of records consumed."""
        broker = self._client._connection
        consumed = 0
        self.stats["steps"] += 1
        for node, sub in self._subs:
            if getattr(node, "_template", None) is not None:
                pipe = self._pipeline(node)
                arena = broker.poll_arena(sub.topics, self._batch_records, self._batch_bytes)
                if arena is None and self._world > 1:
                    arena = _EMPTY_ARENA()                     # lockstep: every rank runs the exchange on every tick
                if arena is not None:
                    consumed += arena.n
                    done = pipe.push(arena)
                    if done is not None:
                        self._produce_fast(node, done)
                elif pipe.pending:                             # input idle (single rank): flush what is in flight
                    for done in pipe.drain():
                        consumed += 1                          # keeps run(until_idle) going: what was just produced may feed a node
                        self._produce_fast(node, done)
                for rb in pipe.take_received():                # forwarded here by the owners of other partitions' inputs
                    self.stats["received"] = self.stats.get("received", 0) + rb.n_publishes
                    broker.produce_publishes(rb)
                continue
            records: list[Record] = broker.poll_batch(sub.topics, self._batch_records, self._batch_bytes)
            if not records:
                continue
            consumed += len(records)
            self._run_host(node, records)
        return consumed

    def close(self) -> None:
        for pipe in self._pipes.values():
            pipe.close()
        self._pipes.clear()

    def _all_idle(self, n: int) -> bool:
        """N > 1: the ranks stop together — a rank with nothing to do still has to run the exchange for the others"""
        if self._world == 1:
            return n == 0
        import torch
        import torch.distributed as dist
        t = torch.tensor([n], dtype=torch.int64, device=torch.device("cuda", self._device))
        dist.all_reduce(t)
        return int(t.item()) == 0

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
            if self._all_idle(n + sum(p.pending_records for p in self._pipes.values())):
                if until_idle:
                    flushed = 0
                    for node, _sub in self._subs:          # nothing in flight on any rank: flush (no exchange involved)
                        pipe = self._pipes.get(id(node))
                        if pipe is not None:
                            for done in pipe.drain():
                                flushed += done.source.n
                                self._produce_fast(node, done)
                            for rb in pipe.take_received():
                                flushed += rb.n_publishes
                                broker.produce_publishes(rb)
                    if self._all_idle(flushed):            # nothing came out that a subscribed node still has to consume
                        return
                    continue
                await asyncio.sleep(idle_sleep)
            else:
                await asyncio.sleep(0)
