"""TEST INFRASTRUCTURE ONLY — times the UNMODIFIED reference's tool-node path (bench.py's CPU arm).

Per event, exactly what the reference's worker does for one consumed record, broker I/O
excluded on both arms (FastStream / aiokafka are not installable offline):

    Envelope.model_validate_json(bytes)                      models/envelope.py:9-17   (FastStream's decode step)
    await node.handler(envelope, correlation_id, broker)     nodes/base.py:149-164 -> nodes/tool.py:37-86 (the sync
                                                             tool runs on the anyio worker thread, as in the stock path)
                                                             -> _publish_action: model_dump_json per publish (capture broker)
    ret.model_dump_json()                                    worker/worker.py:52-53    (handler return -> publish_topic)

One asyncio loop per process, handlers awaited one after the other (max_workers=1, the reference default).
"""
from __future__ import annotations

import asyncio
import os
import sys
import time

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _node():
    from oracle import ref_harness as rh
    ref = rh.load_reference()
    sys.path.insert(0, os.path.join(_ROOT, "tests", "golden"))
    import tools_def
    return rh, ref, ref.agent_tool(tools_def.get_weather)


async def _run(rh, ref, node, chunk, collect):
    out, nbytes = [], 0
    for rec in chunk:
        env = ref.Envelope.model_validate_json(rec)
        corr = env.context.deps.correlation_id
        br = rh.CaptureBroker()
        ret = await node.handler(env, corr, br)
        pubs = br.published
        if node.publish_topic:
            pubs.append((node.publish_topic, None, corr, ret.model_dump_json().encode()))
        for p in pubs:
            nbytes += len(p[3])
        if collect:
            out.append(pubs)
    return out, nbytes


def run_chunk(chunk, collect: bool = False):
    """-> (n events, output bytes, seconds, [publishes per event] if collect)"""
    rh, ref, node = _node()
    t0 = time.perf_counter()
    out, nbytes = asyncio.run(_run(rh, ref, node, chunk, collect))
    return len(chunk), nbytes, time.perf_counter() - t0, out


def _worker(chunk):
    n, nb, dt, _ = run_chunk(chunk)
    return n, nb


class Pool:
    """`cores` worker processes (spawn: the caller may hold a CUDA context), imports and warm-up done once"""
    def __init__(self, cores: int, worker=None):
        import multiprocessing as mp
        self.cores = cores
        self.worker = worker or _worker
        self.pool = mp.get_context("spawn").Pool(cores)

    def run(self, records):
        """-> (events/s, seconds, n)"""
        chunks = [records[i::self.cores] for i in range(self.cores) if records[i::self.cores]]
        t0 = time.perf_counter()
        res = self.pool.map(self.worker, chunks, chunksize=1)
        dt = time.perf_counter() - t0
        n = sum(r[0] for r in res)
        return n / dt, dt, n

    def close(self):
        self.pool.close()
        self.pool.join()


def timed(records, cores: int, start: str = "spawn"):
    """events/s of the reference over `records`, split over `cores` processes -> (events/s, seconds, n)"""
    p = Pool(cores)
    try:
        p.run(records[: max(cores * 8, 64)])                   # imports + warm-up outside the timing
        return p.run(records)
    finally:
        p.close()
