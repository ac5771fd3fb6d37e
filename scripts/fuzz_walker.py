"""Long CPU fuzz of the walker (tests/hostsim build of the device source): every mutant the window-reader walker
accepts must be a fixed point of pydantic's dump(validate(.)), and the window reader and the global reader must
agree on verdict and columns.  usage: python scripts/fuzz_walker.py [n_mutants] [seed]   (test infrastructure)"""
import os, random, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "calfkit-sdk_b200"), ROOT]
from conftest import as_bytes, golden
from hostsim import walk, walk_global
from calfkit import synth
from oracle import port

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
seeds = [as_bytes(c["input"]) for c in golden("codec.json") if c["ok"] and len(as_bytes(c["input"])) < 8000]
seeds += [as_bytes(c["input"]) for c in golden("tool_node.json")]
seeds += synth.tool_events(30, seed=rng.randrange(1 << 30)) + synth.tool_events(10, seed=5, size=None, full_history=True) + \
    synth.mixed_events(20, seed=rng.randrange(1 << 30), hi=6000) + synth.fanout_events(2, seed=6, fanout=9)
seeds = [s for s in seeds if s and walk(s)[0]]
tok = [b'"', b"{", b"}", b"[", b"]", b",", b":", b"\\", b" ", b"\n", b"0", b"1", b"9", b"e", b"E", b".", b"-", b"+", b"null", b"true",
       b"false", b"1.5", b"1e5", b"-0", b"0.10", b'"a"', b"{}", b"[]", b"\\u0041", b"\\/", b"\xc3\xa9", b"\xff", b"\x01", b"\t",
       b'"kind":"tool-return",', b'"a":1,', b"NaN", b"Z", b"+00:00", b".000000", b".5", b"00", b"x" * 40, b'"' + b"y" * 130 + b'"']


def is_fixed(m: bytes) -> bool:
    try:
        return port.encode(port.decode(m)) == m
    except Exception:
        return False


acc = 0
for it in range(n):
    b = bytearray(rng.choice(seeds))
    for _ in range(rng.choice([1, 1, 1, 2, 3])):
        if not b:
            break
        op, i = rng.randrange(7), rng.randrange(len(b))
        if op == 0: b[i] = rng.randrange(256)
        elif op == 1: del b[i]
        elif op == 2: b[i:i] = rng.choice(tok)
        elif op == 3: j = min(len(b), i + rng.randrange(1, 200)); b[i:i] = b[i:j]
        elif op == 4: j = min(len(b), i + rng.randrange(1, 200)); del b[i:j]
        elif op == 5:
            k = bytes(b).find(b"null", i)
            if k >= 0: b[k:k + 4] = rng.choice(tok)
        else:
            k = bytes(b).find(b'"', i)
            if k >= 0: b[k + 1:k + 1] = rng.choice(tok)
    m = bytes(b)
    if not m:
        continue
    a1, c1 = walk(m)
    a2, c2 = walk_global(m)
    assert a1 == a2 and (c1 == c2).all(), ("readers disagree", m[:400])
    if a1:
        acc += 1
        assert is_fixed(m), ("unsound accept", m[:2000])
print(f"{n} mutants, {acc} accepted, all accepted are fixed points; window == global on every mutant")
