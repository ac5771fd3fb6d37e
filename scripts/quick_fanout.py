"""Scratch device-resident timing of the config-3 path (1 Agent -> 64 tools); not a bench."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "calfkit-sdk_b200"))
import torch
from calfkit import synth
from calfkit.engine import BatchEngine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
F = 64
recs = synth.fanout_events(n, seed=3, fanout=F)
b = synth.pack(recs)
e = BatchEngine(0, max_records=n, max_in_bytes=b.data.nbytes + 4096, max_out_bytes=n * (F + 1) * (int(b.data.nbytes / n) + 300) + (1 << 20), max_payloads=n * (F + 1))
reg = {f"tool_{j:02d}": f"tool.tool_{j:02d}.input" for j in range(F)}
e.register_topics(list(reg.values()) + ["planner.input", "planner.output"], num_partitions=8)
e.set_agent_node("planner", "planner.input", "planner.output", reg)
d_in = torch.from_numpy(b.data.copy()).cuda(); d_off = torch.from_numpy(b.offsets.copy()).cuda()
e.profile(True)
for it in range(4):
    e.submit_device(d_in, d_off, n); e.fanout_plan(1767225600000, it, max_fanout=256); e.sync()
    if it == 0: e.profile_read()
for k, (ms, c) in e.profile_read().items():
    if c: print(f"{k:6s} {ms / c:8.3f} ms/launch ({c} launches)")
