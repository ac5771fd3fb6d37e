"""Scratch: device-resident timing of a config-5-like batch (mixed sizes 128 B .. 64 KB, many tools' worth of topics); not a bench."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "calfkit-sdk_b200"))
import numpy as np, torch
from calfkit import synth
from calfkit.engine import BatchEngine, ToolTemplate

n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
t0 = time.time()
base = synth.mixed_events(4096, seed=5, hi=65536, n_agents=256)
recs = [base[i % len(base)] for i in range(n)]
b = synth.pack(recs)
print(f"gen {time.time() - t0:.1f} s, {b.data.nbytes / n:.0f} B/record mean, max {max(len(r) for r in base)} B")
e = BatchEngine(0, max_records=n, max_in_bytes=b.data.nbytes + 4096)
e.register_topics(["tool.get_weather.input", "tool.get_weather.output", "weather_agent.input"] + [f"agent_{i:03d}.input" for i in range(256)], num_partitions=8)
e.set_tool_node("tool.get_weather.output", ToolTemplate.from_format("It's sunny in {location}"))
if os.environ.get("CK_BUCKET", "1") == "1": e.set_bucketing(True)
d_in = torch.from_numpy(b.data.copy()).cuda(); d_off = torch.from_numpy(b.offsets.copy()).cuda()
e.profile(True)
for it in range(4):
    e.submit_device(d_in, d_off, n); e.tool_plan(); e.sync()
    if it == 0: e.profile_read()
prof = e.profile_read()
tot = 0.0
for k, (ms, c) in prof.items():
    if c: print(f"{k:6s} {ms / c:8.3f} ms/launch"); tot += ms / 3
nb, npay, npub = e.out_size()
cols = e.columns()
print(f"sum {tot:.3f} ms per batch of {n} -> {n / tot / 1e3:.1f} M events/s, in+out {(b.data.nbytes + nb) / tot / 1e6:.0f} GB/s; status {np.bincount(cols[0])} actions {np.bincount(cols[1])}")
