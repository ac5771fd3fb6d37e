"""TEST INFRASTRUCTURE ONLY — recipe that makes the UNMODIFIED reference travel to the GPU box.

The reference (calf-ai/calfkit-sdk) is pure Python: there is nothing to compile.  /root/reference exists only in
the build container, so `__graft_entry__.build()` runs this script there: it mirrors the reference's own package
tree, byte for byte, into oracle/_ref/calfkit (git-ignored build output, NOT gpurun-ignored — it ships with the
snapshot like the built .so files) and writes a manifest with the sha256 of every file.  oracle/ref_harness.py then
loads the reference from /root/reference when present and from oracle/_ref otherwise; bench.py's CPU arm
(`--impl reference`, `cpu_baseline.kind == "reference"`) times THAT code.  No reference source is committed.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/calfkit"
DST = os.path.join(HERE, "_ref", "calfkit")


def build(force: bool = False) -> str | None:
    """-> path of the mirrored package, or None when neither the reference nor a previous mirror exists"""
    if not os.path.isdir(SRC):
        return DST if os.path.isdir(DST) else None
    manifest = {}
    for root, _dirs, files in os.walk(SRC):
        for f in files:
            if not (f.endswith(".py") or f.endswith(".typed") or f.endswith(".json")):
                continue
            s = os.path.join(root, f)
            rel = os.path.relpath(s, SRC)
            d = os.path.join(DST, rel)
            os.makedirs(os.path.dirname(d), exist_ok=True)
            data = open(s, "rb").read()
            manifest[rel] = hashlib.sha256(data).hexdigest()
            if force or not os.path.exists(d) or open(d, "rb").read() != data:
                shutil.copyfile(s, d)
    with open(os.path.join(HERE, "_ref", "MANIFEST.json"), "w") as fh:
        json.dump({"source": SRC, "files": manifest}, fh, indent=0, sort_keys=True)
    return DST


if __name__ == "__main__":
    print(build(force=True))
