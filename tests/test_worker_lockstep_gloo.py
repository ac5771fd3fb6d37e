"""World-size-2 gloo test of the partition-sharded Worker's run loop (calfkit/worker/worker.py at N > 1) without GPUs: the lane
pipeline is replaced by a stand-in whose every push performs a collective (as the real exchange does: flag barriers in peer
memory), so a rank that stops ticking while its peer still has work hangs the test instead of passing it.  Checked: ranks
with different amounts of input terminate together, what a rank's step forwards is produced by the OWNER's broker, nothing is
lost or duplicated.  (The idle vote itself — one 8-byte all-reduce — is replaced by its CPU-tensor twin.)"""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank: int, world: int, port: int, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import asyncio
    from calfkit import Client, Worker
    from calfkit.engine._lib import PUB_DTYPE
    from calfkit.engine.lane import Arena, PublishBatch
    from calfkit.nodes import BaseNodeDef

    def batch_of(payloads, topic, source):
        n = len(payloads)
        lens = np.asarray([len(p) for p in payloads], dtype=np.uint32)
        off = np.zeros(n + 1, dtype=np.int64)
        np.cumsum((lens.astype(np.int64) + 15) & ~15, out=off[1:])
        out = np.zeros(int(off[-1]) if n else 0, dtype=np.uint8)
        for i, p in enumerate(payloads):
            out[off[i]:off[i] + len(p)] = np.frombuffer(p, dtype=np.uint8)
        pubs = np.zeros(n, dtype=PUB_DTYPE)
        pubs["payload"], pubs["record"], pubs["topic_id"], pubs["partition"] = np.arange(n), np.arange(n), 0, -1
        return PublishBatch(out, off, lens, pubs, {0: topic}, source, None,
                            None if source is None else np.zeros(source.n, np.uint32), None if source is None else np.zeros(source.n, np.uint32))

    class Node(BaseNodeDef):
        def __init__(self):
            self.node_id, self.subscribe_topics, self.publish_topic = "N", ["t.in"], "t.out"
            self._template = object()

        async def run(self, *a, **k):
            raise NotImplementedError

    class Pipe:
        """a record whose first byte is odd belongs to rank 1, even to rank 0; foreign ones are exchanged at every push"""
        def __init__(self, node):
            self.node, self.inflight, self.received, self.ticks = node, [], [], 0

        def push(self, arena):
            self.ticks += 1
            mine, foreign = [], []
            for i in range(arena.n):
                r = arena.record(i)
                (mine if r[0] % world == rank else foreign).append(r + b"@%d" % rank)
            gathered = [None] * world
            dist.all_gather_object(gathered, foreign)              # the collective of the step: every rank must be here
            got = [p for src, lst in enumerate(gathered) if src != rank for p in lst]
            if got:
                self.received.append(batch_of(got, self.node.publish_topic, None))
            self.inflight.append((arena, mine))
            if len(self.inflight) > 2:
                a, m = self.inflight.pop(0)
                return self._done(a, m)
            return None

        def _done(self, arena, mine):
            # the step's own results: only the payloads this rank owns (the others went to their owner above)
            src = Arena.pack([b"x"] * len(mine)) if mine or arena.n == 0 else Arena.pack([b"x"])
            b = batch_of(mine, self.node.publish_topic, Arena.pack([b"s"] * max(arena.n, 0)) if arena.n else Arena(np.zeros(0, np.uint8), np.zeros(1, np.int64)))
            if arena.n and not mine:
                b.pubs = np.zeros(0, dtype=PUB_DTYPE)
            return b

        def drain(self):
            while self.inflight:
                a, m = self.inflight.pop(0)
                yield self._done(a, m)

        def take_received(self):
            out, self.received = self.received, []
            return out

        @property
        def pending(self):
            return len(self.inflight)

        @property
        def pending_records(self):
            return sum(a.n for a, _m in self.inflight)

        def close(self):
            pass

    def all_idle(self, n):                                          # Worker._all_idle on a CPU tensor (gloo)
        t = torch.tensor([n], dtype=torch.int64)
        dist.all_reduce(t)
        return int(t.item()) == 0

    Worker._all_idle = all_idle
    pipes = {}

    def _pipeline(self, node):                                      # registered where the Worker keeps its pipelines
        if id(node) not in self._pipes:
            self._pipes[id(node)] = pipes[node.node_id] = Pipe(node)
        return self._pipes[id(node)]
    Worker._pipeline = _pipeline
    client = Client.connect()
    broker = client._connection
    seen = []
    broker.sink("t.out", lambda batch, idx: seen.extend(batch.payload(int(j)) for j in idx))
    node = Node()
    worker = Worker(client, nodes=[node], batch_records=4, batch_bytes=1 << 20, rank=rank, world=world)
    # rank 0 has 3 polls' worth of input, rank 1 has one record: they must still stop together
    mine = [bytes([65 + k]) + b"-r%d-%d" % (rank, k) for k in range(11 if rank == 0 else 1)]
    broker.produce_arena("t.in", Arena.pack(mine))
    asyncio.run(worker.run(until_idle=True))
    q.put((rank, mine, sorted(seen), pipes["N"].ticks))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_worker_lockstep_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + (os.getpid() % 90)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(2):
        rank, mine, seen, ticks = q.get(timeout=180)
        res[rank] = (mine, seen, ticks)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][2] == res[1][2] >= 5                               # the same number of ticks on both ranks: 3 polls + 2 to empty the pipe
    everything = [r + b"@%d" % rk for rk in (0, 1) for r in res[rk][0]]
    for rk in (0, 1):
        assert res[rk][1] == sorted(p for p in everything if p[0] % 2 == rk)      # each payload produced once, by its owner
