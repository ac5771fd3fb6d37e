"""Per-source-line instruction / stall-sample totals from a .ncu-rep captured with --import-source on.
usage: python scripts/ncu_lines.py rep.ncu-rep [top]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = None; hdr = None
tot = collections.defaultdict(lambda: [0, 0, ""])
for r in csv.reader(out.splitlines()):
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; ie = hdr.index("Instructions Executed"); sm = hdr.index("# Samples"); continue
    if hdr and r[0].isdigit() and r[2] == "-":
        try:
            k = (cur, int(r[0])); tot[k][0] += int(r[ie]); tot[k][1] += int(r[sm]); tot[k][2] = r[1].strip()[:90]
        except ValueError: pass
ti = sum(v[0] for v in tot.values()); ts = sum(v[1] for v in tot.values())
print("total warp-instr", ti, "samples", ts)
print("--- by instructions")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*v[0]/ti:5.1f}% inst {100*v[1]/max(ts,1):5.1f}% smp  {k[0]}:{k[1]:<5d} {v[2]}")
print("--- by stall samples")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{100*v[0]/ti:5.1f}% inst {100*v[1]/max(ts,1):5.1f}% smp  {k[0]}:{k[1]:<5d} {v[2]}")
