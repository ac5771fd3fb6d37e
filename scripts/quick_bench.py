"""Scratch device-resident timing of the tool-node path (not the contract bench: see bench.py)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "calfkit-sdk_b200"))
import numpy as np, torch
from calfkit import synth
from calfkit.engine import BatchEngine, ToolTemplate

n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
t0 = time.time(); recs = synth.tool_events(n, seed=0); b = synth.pack(recs); print("gen", round(time.time() - t0, 1), "s", b.data.nbytes / n, "B/rec")
e = BatchEngine(0, max_records=n, max_in_bytes=b.data.nbytes + 4096)
e.register_topics(["tool.get_weather.input", "tool.get_weather.output", "weather_agent.input"], num_partitions=8)
e.set_tool_node("tool.get_weather.output", ToolTemplate.from_format("It's sunny in {location}"))
d_in = torch.from_numpy(b.data).cuda(); d_off = torch.from_numpy(b.offsets).cuda()
torch.cuda.synchronize()
e.profile(True)
for it in range(4):
    e.submit_device(d_in, d_off, n); e.tool_plan(); e.sync()
    if it == 0: e.profile_read()
prof = e.profile_read()
nb, npay, npub = e.out_size()
print("out bytes", nb, "payloads", npay, "pubs", npub)
for k, (ms, cnt) in prof.items():
    if cnt: print(f"{k:6s} {ms / cnt:8.3f} ms/launch  ({cnt} launches)")
tot = sum(ms for ms, c in prof.values()) / 3
print("sum of kernels per batch %.3f ms -> %.1f M events/s ; in+out GB/s %.1f" % (tot, n / tot / 1e3, (b.data.nbytes + nb) / tot / 1e6))
cols = e.columns(); print("status counts", np.bincount(cols[0]), "actions", np.bincount(cols[1]))
e.profile(False)
st = torch.cuda.ExternalStream(e.stream_ptr())
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(st):
    ev0.record(st)
    for it in range(5):
        e.submit_device(d_in, d_off, n); e.tool_plan()
    ev1.record(st)
e.sync(); torch.cuda.synchronize()
ms = ev0.elapsed_time(ev1) / 5
print("pipelined: %.3f ms/batch -> %.1f M events/s" % (ms, n / ms / 1e3))
