"""2-GPU diagnosis of the exchange cost (scratch): compute only / exchange only / alternating lanes."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "calfkit-sdk_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
import bench
from calfkit import synth
from calfkit.engine import BatchEngine, ToolTemplate
from calfkit.engine._lib import COL
from calfkit.engine.exchange import PeerExchange
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
opts = dist.ProcessGroupNCCL.Options(); opts.is_high_priority_stream = True
dist.init_process_group("nccl", device_id=dev, pg_options=opts)
n = 1_000_000
batch = synth.pack(synth.tool_events(n, seed=1000 + rank))
topics = ["tool.get_weather.input", "tool.get_weather.output", "weather_agent.input"]
def mk():
    e = BatchEngine(local, max_records=n, max_in_bytes=batch.data.nbytes + 4096)
    e.register_topics(topics, num_partitions=8); e.set_tool_node("tool.get_weather.output", ToolTemplate.from_format("It's sunny in {location}"))
    fwd = int(n * 0.3) + 1024
    return e, PeerExchange(e, rank, world, max_fwd=fwd, data_cap=fwd * 1300)
lanes = [mk(), mk()]
e0 = lanes[0][0]
e0.submit(batch.data, batch.offsets)
corr_off = e0.columns()[COL["CORR_OFF"]]
batch = bench.place_on_partitions(batch, corr_off, rank, world, 0.125, seed=2000 + rank)
d_in = torch.from_numpy(batch.data.copy()).to(dev); d_off = torch.from_numpy(batch.offsets.copy()).to(dev)
def compute(l): l[0].submit_device(d_in, d_off, n); l[0].tool_plan()
step = [0]
def exch(l): step[0] += 1; l[1].send(step[0])
def timed(fn, reps=10):
    dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter(); fn(reps); torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / reps * 1e3
    t = torch.tensor([dt], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); return float(t.item())
for l in lanes: compute(l); exch(l)
torch.cuda.synchronize()
r = {}
r["compute_only"] = timed(lambda k: [compute(lanes[0]) for _ in range(k)])
r["compute_then_exchange_one_lane"] = timed(lambda k: [(compute(lanes[0]), exch(lanes[0])) for _ in range(k)])
def only_x(k):
    for _ in range(k): exch(lanes[0])
r["exchange_only"] = timed(only_x)
lanes[0][0].profile(True)
only_x(5); torch.cuda.synchronize()
r["exchange_kernel_ms"] = {k: round(ms / max(c, 1), 3) for k, (ms, c) in lanes[0][0].profile_read().items() if c}
lanes[0][0].profile(False)
def alt(k):
    compute(lanes[0])
    for i in range(1, k): compute(lanes[i % 2]); exch(lanes[(i - 1) % 2])
    exch(lanes[(k - 1) % 2])
r["alternating_lanes"] = timed(alt)
if rank == 0: print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in r.items()}, flush=True)
for e, px in lanes:
    del px
import gc; lanes_e = [e for e, _ in lanes]; lanes = None; gc.collect(); torch.cuda.synchronize()
for e in lanes_e: e.close()
dist.barrier(); dist.destroy_process_group()
