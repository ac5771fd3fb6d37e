"""Long CPU fuzz of the canonicaliser + decode pass (tests/hostsim build of the device sources) against pydantic: verdict
(ok / json_invalid / schema-invalid / declared unsupported) and canonical bytes must agree on every mutant that is not
declared unsupported, and the trusting second walk must accept every re-emission.  usage: python scripts/fuzz_canon.py [n] [seed]"""
import json, os, random, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "calfkit-sdk_b200"), ROOT]
import test_canon_hostsim as T
from conftest import as_bytes, golden
from calfkit import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
seeds = [as_bytes(c["input"]) for c in golden("codec.json") + golden("tool_node.json") if 0 < len(as_bytes(c["input"])) < 5000]
seeds += synth.tool_events(6, seed=rng.randrange(1 << 30)) + synth.tool_events(4, seed=8, size=None, full_history=True)
seeds += [c["input"].encode() for c in golden("replies.json")] + [c["input"].encode() for c in golden("agent_run.json")]
for s in list(seeds):
    try:
        obj = json.loads(s)
    except Exception:
        continue
    seeds.append(json.dumps(obj, indent=1).encode())
    seeds.append(json.dumps(obj, sort_keys=True, ensure_ascii=True, separators=(", ", " : ")).encode())
tok = [b'"', b"{", b"}", b"[", b"]", b",", b":", b"\\", b" ", b"\n", b"0", b"1", b"9", b"e", b"E", b".", b"-", b"+", b"null", b"true", b"false",
       b"1.5", b"1e5", b"-0", b"0.10", b'"a"', b"{}", b"[]", b"\\u0041", b"\\/", b"\xc3\xa9", b"\xff", b"\x01", b"\t", b'"kind":"tool-return",',
       b'"a":1,', b"NaN", b"Infinity", b"Z", b"+00:00", b"+0530", b".000000", b".5", b",5", b"00", b'"zz":[1,{"q":2}],', b'"part_kind":"text",',
       b'"kind":"request",', b"\\ud83d\\ude00", b"\\ud800", b"2.50", b"1E-7", b"12e3", b"0.30000000000000004", b"21.700000000000003",
       b"1.2345678901234567e+30", b"0.1234567890123456789", b" ", b"_"]
stats: dict = {}
for _ in range(n):
    b = bytearray(rng.choice(seeds))
    for _ in range(rng.choice([1, 1, 1, 2, 3])):
        if not b:
            break
        op, i = rng.randrange(7), rng.randrange(len(b))
        if op == 0: b[i] = rng.randrange(256)
        elif op == 1: del b[i]
        elif op == 2: b[i:i] = rng.choice(tok)
        elif op == 3: j = min(len(b), i + rng.randrange(1, 40)); b[i:i] = b[i:j]
        elif op == 4: j = min(len(b), i + rng.randrange(1, 40)); del b[i:j]
        elif op == 5:
            k = bytes(b).find(b"null", i)
            if k >= 0: b[k:k + 4] = rng.choice(tok)
        else:
            k = bytes(b).find(b'"', i)
            if k >= 0: b[k + 1:k + 1] = rng.choice(tok)
    if b:
        T._check(bytes(b), stats)
print(f"{n} mutants: verdicts {dict(sorted(stats.items()))} (0 ok, 2 json_invalid, 3 schema-invalid, 4 declared unsupported); every ok byte-identical to pydantic")
