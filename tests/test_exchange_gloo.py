"""World-size-2 gloo test of the cross-partition exchange logic (calfkit/engine/exchange.py): the same
planning code the NCCL path uses, on CPU tensors, with a torch gather standing in for ck_gather_spans."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank: int, world: int, port: int, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from calfkit.engine.exchange import exchange, plan_exchange
    rng = np.random.default_rng(rank)
    n = 50
    lens = rng.integers(5, 40, size=n)
    out_off = torch.from_numpy(np.concatenate([[0], np.cumsum(lens)]).astype(np.int64))
    out = torch.from_numpy(rng.integers(0, 255, size=int(out_off[-1]), dtype=np.uint8))
    part = rng.integers(0, 8, size=n)
    pubs = torch.zeros((2 * n, 8), dtype=torch.int32)
    pubs[:, 0] = -1
    for i in range(n):                      # slot 2i: keyed callback publish, 2i+1: unkeyed publish_topic copy
        pubs[2 * i] = torch.tensor([i, 2, 0, 0, i, 1, int(part[i]), 0], dtype=torch.int32)
        pubs[2 * i + 1] = torch.tensor([i, 1, 0, 0, i, 0, -1, 0], dtype=torch.int32)
    plan = plan_exchange(pubs, out_off, torch.from_numpy(lens.astype(np.int32)), rank, world)
    expect_sel = [2 * i for d in range(world) for i in range(n) if part[i] % world == d and d != rank]
    assert plan.sel.tolist() == expect_sel
    send, recv = torch.zeros(4096, dtype=torch.uint8), torch.zeros(4096, dtype=torch.uint8)

    def gather(p, buf):
        for s, l, d in zip(p.src_off.tolist(), p.lens.tolist(), p.dst_off.tolist()):
            buf[d:d + l] = out[s:s + l]
    nrecv, rbytes, rlens = exchange(plan, gather, send, recv)
    sent = [bytes(out[int(out_off[i]):int(out_off[i + 1])].tolist()) for i in range(n) if part[i] % world != rank]
    got, p = [], 0
    for l in rlens.tolist():
        got.append(bytes(recv[p:p + l].tolist()))
        p += l
    assert p == rbytes and len(got) == nrecv
    q.put((rank, sent, got))
    dist.barrier()
    dist.destroy_process_group()


def test_exchange_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 400)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(2):
        rank, sent, got = q.get(timeout=120)
        res[rank] = (sent, got)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][0] == res[1][1] and res[1][0] == res[0][1]      # what one rank sent is what the other received
    assert len(res[0][0]) > 0 and len(res[1][0]) > 0
