"""N-rank check of the RECEIVED bytes of the cross-partition forward (run under torchrun on N GPUs of one box):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/multigpu/exchange_parity.py

Every rank builds every rank's seeded batch, so it knows — through the oracle port — exactly which payloads the others must
forward to it (keyed publishes whose murmur2(correlation_id) % 8 % world is this rank), in which order and with which topic;
the regions the peers' kernels wrote into this rank's receive buffer must hold those bytes.  Prints one JSON line per rank."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (os.path.join(ROOT, "calfkit-sdk_b200"), ROOT, os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)


def murmur2(data: bytes) -> int:
    m, h = 0x5BD1E995, (0x9747B28C ^ len(data)) & 0xFFFFFFFF
    n4 = len(data) // 4
    for i in range(n4):
        k = int.from_bytes(data[4 * i:4 * i + 4], "little")
        k = (k * m) & 0xFFFFFFFF; k ^= k >> 24; k = (k * m) & 0xFFFFFFFF
        h = (h * m) & 0xFFFFFFFF; h ^= k
    t = data[4 * n4:]
    if len(t) == 3: h ^= t[2] << 16
    if len(t) >= 2: h ^= t[1] << 8
    if len(t) >= 1: h ^= t[0]; h = (h * m) & 0xFFFFFFFF
    h ^= h >> 13; h = (h * m) & 0xFFFFFFFF; h ^= h >> 15
    return h


def main():
    import numpy as np
    import torch
    import torch.distributed as dist
    import tools_def
    from calfkit import synth
    from calfkit.engine import BatchEngine, ToolTemplate
    from calfkit.engine.exchange import PeerExchange
    from oracle import port
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n, P = 3000, 8
    batches = [synth.tool_events(n, seed=500 + r) + synth.tool_events(64, seed=600 + r, size=None, full_history=True) for r in range(world)]
    topics = ["tool.get_weather.input", "tool.get_weather.output", "weather_agent.input"]
    eng = BatchEngine(local, max_records=4096, max_in_bytes=16 << 20)
    eng.register_topics(topics, num_partitions=P)
    eng.set_tool_node("tool.get_weather.output", ToolTemplate.from_format("It's sunny in {location}"))
    px = PeerExchange(eng, rank, world, max_fwd=4096, data_cap=8 << 20)
    node = port.ToolNode.of(tools_def.get_weather)
    ok, checked = True, 0
    for step in (1, 2, 3):                                   # several steps: regions are reused, barriers order them
        recs = [r if step % 2 else r for r in batches[rank]]
        b = synth.pack(recs)
        eng.submit(b.data, b.offsets)
        eng.tool_plan()
        px.send(step)
        got = px.received(step)
        for src, meta, data in got:
            want = []
            for r in batches[src]:
                for (tp, key, _c, payload) in port.tool_node_event(node, r):
                    if key is not None and (murmur2(key) & 0x7FFFFFFF) % P % world == rank:
                        want.append((tp, (murmur2(key) & 0x7FFFFFFF) % P, payload))
            have = list(zip([eng.topic_names[int(t)] for t in meta["topic_id"]], [int(p) for p in meta["partition"]], PeerExchange.payloads(meta, data)))
            ok = ok and have == want
            checked += len(want)
    res = torch.tensor([1 if ok else 0, checked], device="cuda")
    dist.all_reduce(res, op=dist.ReduceOp.MIN)
    print(json.dumps({"rank": rank, "world": world, "exchange_parity": bool(ok), "payloads_checked": checked, "all_ranks_ok": bool(res[0].item())}), flush=True)
    eng.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
