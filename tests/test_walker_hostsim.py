"""CPU fuzz of the device walker source (g++ build, tests/hostsim): it must accept ONLY byte strings
that are fixed points of the reference codec (soundness), and must accept every workload shape the
benchmarks use (completeness on those shapes)."""
import random

from conftest import as_bytes, golden
from hostsim import walk
from pydantic import ValidationError


def _is_fixed(b: bytes) -> bool:
    from calfkit.models import Envelope
    try:
        return Envelope.model_validate_json(b).model_dump_json().encode() == b
    except ValidationError:
        return False


def test_goldens_sound_and_expected_rejects():
    conservative = {"any_canonical_numbers", "depth_100", "depth_190", "tool_result_tagged_invalid_falls_to_any"}
    for c in golden("codec.json"):
        b = as_bytes(c["input"])
        acc, _ = walk(b)
        fixed = c["ok"] and c["output"].encode() == b
        assert not (acc and not fixed), c["name"]
        if fixed and c["name"] not in conservative:
            assert acc, c["name"]


def test_accepts_every_synthetic_shape():
    from calfkit import synth
    recs = (synth.tool_events(300, seed=1) + synth.tool_events(100, seed=2, size=None, full_history=True)
            + synth.fanout_events(4, seed=3, fanout=64) + synth.mixed_events(60, seed=4, hi=40000))
    for r in recs:
        acc, cols = walk(r)
        assert acc, (cols[2], r[max(0, int(cols[2]) - 40):int(cols[2]) + 40])


def test_mutation_fuzz_is_sound():
    from calfkit import synth
    rng = random.Random(7)
    seeds = [as_bytes(c["input"]) for c in golden("codec.json")
             if c["ok"] and c["output"].encode() == as_bytes(c["input"]) and len(as_bytes(c["input"])) < 6000]
    seeds += synth.tool_events(20, seed=9) + synth.tool_events(10, seed=8, size=None, full_history=True)
    seeds += [as_bytes(c["input"]) for c in golden("tool_node.json")
              if c["name"] in ("frame_overrides", "any_values", "existing_same_result_last", "three_frames", "wf_metadata")]
    seeds = [s for s in seeds if walk(s)[0]]
    tok = [b'"', b"{", b"}", b"[", b"]", b",", b":", b"\\", b" ", b"\n", b"0", b"1", b"9", b"e", b"E", b".", b"-", b"+",
           b"null", b"true", b"false", b"1.5", b"1e5", b"-0", b"0.10", b'"a"', b"{}", b"[]", b"\\u0041", b"\\/",
           b"\xc3\xa9", b"\xff", b"\x01", b"\t", b'"kind":"tool-return",', b'"a":1,', b"NaN", b"Z", b"+00:00",
           b".000000", b".5", b"00"]
    accepted = 0
    for _ in range(120000):
        b = bytearray(rng.choice(seeds))
        for _ in range(rng.choice([1, 1, 1, 2, 3])):
            if not b:
                break
            op, i = rng.randrange(7), rng.randrange(len(b))
            if op == 0:
                b[i] = rng.randrange(256)
            elif op == 1:
                del b[i]
            elif op == 2:
                b[i:i] = rng.choice(tok)
            elif op == 3:
                j = min(len(b), i + rng.randrange(1, 40)); b[i:i] = b[i:j]
            elif op == 4:
                j = min(len(b), i + rng.randrange(1, 40)); del b[i:j]
            elif op == 5:
                k = bytes(b).find(b"null", i)
                if k >= 0:
                    b[k:k + 4] = rng.choice(tok)
            else:
                k = bytes(b).find(b'"', i)
                if k >= 0:
                    b[k + 1:k + 1] = rng.choice(tok)
        m = bytes(b)
        if m and walk(m)[0]:
            accepted += 1
            assert _is_fixed(m), m[:2000]
    assert accepted > 1000


def test_table_driven_walker_equals_recursive_walker():
    """csrc/ck_vm.cuh (bytecode interpreter, the one the GPU runs) against csrc/ck_walk.cuh (recursive
    descent, the one fuzzed against pydantic above): same verdict and same columns on every input."""
    from calfkit import synth
    from hostsim import vm_walk, walk_global
    rng = random.Random(11)
    seeds = [as_bytes(c["input"]) for c in golden("codec.json")] + [as_bytes(c["input"]) for c in golden("tool_node.json")]
    seeds += synth.tool_events(30, seed=1) + synth.tool_events(10, seed=2, size=None, full_history=True) + \
        synth.fanout_events(2, seed=3, fanout=70) + synth.mixed_events(20, seed=4, hi=20000)
    seeds = [s for s in seeds if s]

    def same(b):
        a1, c1 = walk(b)
        a2, c2 = vm_walk(b)
        a3, c3 = walk_global(b)                  # window reader (product) vs plain global reader: identical in every column
        assert a1 == a2 == a3, b[:300]
        assert (c1 == c3).all(), b[:300]
        if a1:
            assert (c1[2:50] == c2[2:50]).all(), b[:300]
    for s in seeds:
        same(s)
    good = [s for s in seeds if walk(s)[0]]
    tok = [b'"', b"{", b"}", b"[", b"]", b",", b":", b"\\", b" ", b"0", b"1", b"e", b".", b"-", b"null", b"true", b"false",
           b"1.5", b'"a"', b"{}", b"[]", b"\xc3\xa9", b"\xff", b"\x01", b'"kind":"tool-return",', b'"a":1,', b"Z", b".5"]
    for _ in range(60000):
        b = bytearray(rng.choice(good))
        for _ in range(rng.choice([1, 1, 2, 3])):
            if not b:
                break
            op, i = rng.randrange(5), rng.randrange(len(b))
            if op == 0:
                b[i] = rng.randrange(256)
            elif op == 1:
                del b[i]
            elif op == 2:
                b[i:i] = rng.choice(tok)
            elif op == 3:
                j = min(len(b), i + rng.randrange(1, 40)); b[i:i] = b[i:j]
            else:
                j = min(len(b), i + rng.randrange(1, 40)); del b[i:j]
        if b:
            same(bytes(b))
