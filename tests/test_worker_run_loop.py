"""Worker.run(until_idle=True) over pipelined lanes, without a GPU: the lane pipeline is replaced by a stand-in with the same
surface (push returns the results of an OLDER step once the pipe is full, drain flushes, take_received hands over what peers
forwarded), so what is tested is the run loop itself — in particular that it keeps going while a flush has produced records a
subscribed node still has to consume (a regression of this kind reached the GPU suite once)."""
import asyncio

import numpy as np
import pytest

from calfkit import Client, Worker
from calfkit.engine._lib import PUB_DTYPE
from calfkit.engine.lane import Arena, PublishBatch
from calfkit.nodes import BaseNodeDef


class TemplateNode(BaseNodeDef):
    def __init__(self, name, sub, pub):
        self.node_id, self.subscribe_topics, self.publish_topic = name, [sub], pub
        self._template = object()          # what makes Worker.step take the lane path

    async def run(self, *a, **k):
        raise NotImplementedError


class FakePipe:
    """K = 3 lanes: a push returns the results of the step pushed two calls earlier"""
    def __init__(self, node, topic_id=0):
        self.node, self.inflight, self.pushed = node, [], 0

    def _result(self, arena):
        n = arena.n
        payloads = [arena.record(i) + b"+" + self.node.node_id.encode() for i in range(n)]
        lens = np.asarray([len(p) for p in payloads], dtype=np.uint32)
        off = np.zeros(n + 1, dtype=np.int64)
        np.cumsum((lens.astype(np.int64) + 15) & ~15, out=off[1:])
        out = np.zeros(int(off[-1]) if n else 0, dtype=np.uint8)
        for i, p in enumerate(payloads):
            out[off[i]:off[i] + len(p)] = np.frombuffer(p, dtype=np.uint8)
        pubs = np.zeros(n, dtype=PUB_DTYPE)
        pubs["payload"], pubs["record"], pubs["topic_id"], pubs["partition"] = np.arange(n), np.arange(n), 0, -1
        return PublishBatch(out, off, lens, pubs, {0: self.node.publish_topic}, arena, None, np.zeros(n, np.uint32), np.zeros(n, np.uint32),
                            on_release=arena.release)

    def push(self, arena):
        self.pushed += 1
        self.inflight.append(arena)
        return self._result(self.inflight.pop(0)) if len(self.inflight) > 2 else None

    def drain(self):
        while self.inflight:
            yield self._result(self.inflight.pop(0))

    def take_received(self):
        return []

    @property
    def pending(self):
        return len(self.inflight)

    @property
    def pending_records(self):
        return sum(a.n for a in self.inflight)

    def close(self):
        pass


@pytest.fixture()
def rig(monkeypatch):
    client = Client.connect()
    a = TemplateNode("A", "t.a", "t.b")
    b = TemplateNode("B", "t.b", "t.c")
    pipes = {}
    def _pipeline(self, node):                          # registered where the Worker keeps its pipelines (the final flush looks there)
        if id(node) not in self._pipes:
            self._pipes[id(node)] = pipes[node.node_id] = FakePipe(node)
        return self._pipes[id(node)]
    monkeypatch.setattr(Worker, "_pipeline", _pipeline)
    return client, client._connection, a, b, pipes


def test_chain_of_two_template_nodes_drains_completely(rig):
    client, broker, a, b, pipes = rig
    worker = Worker(client, nodes=[a, b], batch_records=4, batch_bytes=1 << 20)
    vals = [f"v{i}".encode() for i in range(10)]
    broker.produce_arena("t.a", Arena.pack(vals))
    asyncio.run(worker.run(until_idle=True))
    got = sorted(r.value for r in broker.poll_batch(("t.c",), 100)) if "t.c" in broker.subscribed else None
    # t.c has no subscriber: its publishes are counted as dropped; A's and B's work is visible in the counters
    assert got is None and broker.dropped_unsubscribed == 10
    assert worker.stats["records"] == 20 and worker.stats["publishes"] == 20     # every record went through A and then through B
    assert pipes["A"].pushed == 3 and pipes["B"].pushed >= 3                     # 10 records in polls of 4
    assert broker.pending() == 0 and all(p.pending == 0 for p in pipes.values())


@pytest.mark.parametrize("consumer_first", [False, True])
def test_outputs_reach_a_sink_in_order(rig, consumer_first):
    """consumer_first: B is visited before A in a step, so what A's flush produces is only seen by the NEXT step — the loop
    must not stop in between"""
    client, broker, a, b, pipes = rig
    seen = []
    broker.sink("t.c", lambda batch, idx: seen.extend(batch.payload(int(j)) for j in idx))
    worker = Worker(client, nodes=[b, a] if consumer_first else [a, b], batch_records=3, batch_bytes=1 << 20)
    vals = [f"v{i}".encode() for i in range(8)]
    broker.produce_arena("t.a", Arena.pack(vals))
    asyncio.run(worker.run(until_idle=True))
    assert seen == [v + b"+A+B" for v in vals]
    # fewer records than the pipe is deep: A's only output comes from the flush of an idle step
    del seen[:]
    broker.produce_arena("t.a", Arena.pack(vals[:2]))
    asyncio.run(worker.run(until_idle=True))
    assert seen == [v + b"+A+B" for v in vals[:2]]


def test_idle_worker_returns_at_once(rig):
    client, broker, a, b, pipes = rig
    worker = Worker(client, nodes=[a, b])
    asyncio.run(worker.run(until_idle=True))
    assert worker.stats["records"] == 0
