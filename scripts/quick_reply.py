"""Scratch: walk time on reply-shaped batches, one shape per batch vs shapes mixed inside every warp (not a bench)."""
import os, sys, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "calfkit-sdk_b200"))
import torch
from calfkit import synth
from calfkit.engine import BatchEngine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
base = synth.tool_events(20000, seed=1)
TEXT = '{"kind":"text","text":"It\'s sunny in %s","metadata":null}'
DATA = '{"kind":"data","data":{"temp":%d,"ok":true,"tags":["a","b"]},"schema_":null,"metadata":null}'


def make(kinds, sort_by_kind=False):
    rng = random.Random(3)
    ks = [rng.choice(kinds) for _ in range(n)]
    if sort_by_kind:
        ks.sort()
    out = []
    for i, k in enumerate(ks):
        parts = [] if k == 0 else [TEXT % ("x" * rng.randrange(4, 40))]
        if k >= 2:
            parts.insert(k - 2, DATA % rng.randrange(40))
        out.append(base[i % len(base)].replace(b'"final_output_parts":[]', ('"final_output_parts":[' + ",".join(parts) + "]").encode()))
    return out


for label, kinds, srt in (("shape 0 only", [0], False), ("shape 1 only", [1], False), ("shape 3 only", [3], False),
                          ("4 shapes mixed", [0, 1, 2, 3], False), ("4 shapes, bucketed by shape", [0, 1, 2, 3], True)):
    b = synth.pack(make(kinds, srt))
    e = BatchEngine(0, max_records=n, max_in_bytes=b.data.nbytes + 4096)
    d_in = torch.from_numpy(b.data.copy()).cuda(); d_off = torch.from_numpy(b.offsets.copy()).cuda()
    e.profile(True)
    for it in range(4):
        e.submit_device(d_in, d_off, n); e.reply_plan(0); e.sync()
        if it == 0: e.profile_read()
    prof = e.profile_read()
    print(f"{label:32s} walk {prof['walk'][0] / prof['walk'][1]:.3f} ms per {n} records ({b.data.nbytes / n:.0f} B/record)")
    e.close()
