"""Partition-sharded Worker on N GPUs (run under torchrun): every rank runs calfkit.Worker.run() over its own shard of the
input; keyed publishes whose partition another rank owns are forwarded over NVLink and produced THERE.  Checked against the
oracle over the union of all ranks' inputs: the multiset of (topic, payload) produced anywhere, and that every keyed publish
was produced by the rank that owns its partition.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 tests/multigpu/worker_sharded.py"""
import asyncio
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (os.path.join(ROOT, "calfkit-sdk_b200"), ROOT, os.path.join(ROOT, "tests", "golden"), os.path.join(ROOT, "tests", "multigpu")):
    sys.path.insert(0, p)


def main():
    import torch
    import torch.distributed as dist
    import tools_def
    from calfkit import Client, Worker, agent_tool, synth
    from calfkit.engine.lane import Arena
    from exchange_parity import murmur2
    from oracle import port
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    P = 8
    batches = [synth.tool_events(2500 + 300 * r, seed=700 + r) + synth.tool_events(40, seed=800 + r, size=None, full_history=True) for r in range(world)]
    node_def = agent_tool(tools_def.get_weather, device_template="It's sunny in {location}")
    client = Client.connect("localhost")
    client.broker.num_partitions = P
    mine = batches[rank]
    for a in range(0, len(mine), 700):                         # several polls per rank, and a different number per rank
        client.broker.produce_arena("tool.get_weather.input", Arena.pack(mine[a:a + 700]))
    got = []
    for t in ("weather_agent.input", "tool.get_weather.output"):
        client.broker.sink(t, lambda b, idx, t=t: got.extend((t, int(b.pubs["partition"][j]), p) for (_t, _k, p, j) in b.iter_records(idx)))
    worker = Worker(client, nodes=[node_def], device=local, batch_records=1024, batch_bytes=2 << 20, lanes=3,
                    route_topics=["weather_agent.input"], rank=rank, world=world)
    asyncio.run(worker.run(until_idle=True))
    # ownership: a keyed publish (the callback) must have been produced by the owner of its partition
    own_ok = all(part % world == rank for (t, part, _p) in got if t == "weather_agent.input")
    allg = [None] * world
    dist.all_gather_object(allg, [(t, p) for (t, _part, p) in got])
    node = port.ToolNode.of(tools_def.get_weather)
    want = sorted((tp, pl) for b in batches for r in b for (tp, _k, _c, pl) in port.tool_node_event(node, r))
    have = sorted(x for g in allg for x in g)
    # and the partition each callback landed on is the key's
    part_ok = True
    for (t, part, p) in got:
        if t == "weather_agent.input":
            corr = json.loads(p)["context"]["deps"]["correlation_id"].encode()
            part_ok = part_ok and part == (murmur2(corr) & 0x7FFFFFFF) % P
    ok = own_ok and part_ok and have == want
    res = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(res, op=dist.ReduceOp.MIN)
    print(json.dumps({"rank": rank, "world": world, "worker_sharded_parity": bool(ok), "produced_here": len(got), "received_from_peers": worker.stats.get("received", 0),
                      "union_matches_oracle": have == want, "all_ranks_ok": bool(res.item())}), flush=True)
    worker.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
