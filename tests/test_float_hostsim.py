"""16-17 digit floats (csrc/ck_float.cuh): the exact "literal == repr(its double)" test against Python, and the
walker / canonicaliser using it against pydantic, on the CPU build of the device sources (tests/hostsim)."""
import ctypes
import json
import math
import random
import struct

from conftest import as_bytes, golden  # noqa: F401
from hostsim import canon, decode, lib, walk


def _digits(x: float):
    s = repr(x)
    mant, _, ex = s.partition("e")
    ip, _, fp = mant.partition(".")
    ds = (ip + fp).lstrip("0")
    k = (int(ex) if ex else 0) - len(fp)
    while ds.endswith("0"):
        ds, k = ds[:-1], k + 1
    return int(ds), k


def _is_repr(m: int, k: int) -> bool:
    L = lib()
    L.ck_host_is_repr.argtypes = [ctypes.c_uint64, ctypes.c_int]
    L.ck_host_is_repr.restype = ctypes.c_int
    return bool(L.ck_host_is_repr(m, k))


def test_exact_repr_test_agrees_with_python():
    rng = random.Random(5)
    n_long = 0
    for _ in range(60000):
        mode = rng.randrange(4)
        if mode == 0:
            x = rng.random() * 10 ** rng.randrange(-6, 17)
        elif mode == 1:
            x = struct.unpack("<d", struct.pack("<Q", rng.getrandbits(63)))[0]
        elif mode == 2:
            x = rng.uniform(0, 100) + rng.uniform(0, 1e-12)
        else:
            x = 2.0 ** rng.randrange(-60, 60) * (1 + rng.choice([0, 2 ** -52, -2 ** -53, rng.random()]))
        if not math.isfinite(x) or not (1e-290 < x < 1e290):
            continue
        m, k = _digits(x)
        if len(str(m)) >= 16:
            n_long += 1
            assert _is_repr(m, k), repr(x)
            for dm in (-2, -1, 1, 2, 9):                          # nearby literals: accepted only if they are repr of their own double
                m2 = m + dm
                if m2 % 10 == 0 or len(str(m2)) != len(str(m)):
                    continue
                y = float(f"{m2}e{k}")
                assert _is_repr(m2, k) == (_digits(y) == (m2, k)), (m2, k, repr(y))
    assert n_long > 20000


def _with_values(values_json: str) -> bytes:
    from calfkit import synth
    r = synth.tool_events(1, seed=3)[0]
    i = r.index(b'"provided_deps":{') + len(b'"provided_deps":{')
    j = r.index(b"}", i)
    return r[:i] + b'"v":' + values_json.encode() + r[j:]


def test_decode_pass_accepts_long_floats_iff_valid():
    """first walk leaves 16-17 digit floats to the canonicaliser (exact test), whose output the second walk trusts:
    the pass as a whole must turn every such record into exactly pydantic's canonical bytes"""
    from oracle import port
    rng = random.Random(9)
    n_ok = n_fixed = 0
    for _ in range(3000):
        vals = []
        for _ in range(rng.randrange(1, 6)):
            x = rng.choice([rng.random(), 0.1 + 0.2, rng.uniform(-1e6, 1e6), rng.random() * 1e-7, rng.random() * 1e18, 2.0 ** rng.randrange(-40, 40) / 3])
            s = repr(x)
            if rng.random() < 0.3:                                   # perturb the last digit: usually no longer the shortest spelling
                mant, e, ex = s.partition("e")
                if mant[-1].isdigit() and mant[-1] not in "09":
                    mant = mant[:-1] + str(int(mant[-1]) + rng.choice([-1, 1]))
                s = mant + e + ex
            vals.append(s)
        rec = _with_values("[" + ",".join(vals) + "]")
        want = port.encode(port.decode(rec))
        assert not walk(rec)[0] or want == rec                       # the first walk stays sound
        st, out = decode(rec)
        assert st in (0, 4), st                                      # 4 = declared unsupported (a literal that is not the shortest spelling)
        if st == 0:
            assert out == want, (vals, out[:200])
            n_ok += 1
        if want == rec:
            n_fixed += 1
            assert st == 0, ("fixed point not accepted", vals)
    assert n_ok > 500 and n_fixed > 300


def test_canonicaliser_keeps_long_floats():
    from oracle import port
    rng = random.Random(10)
    n_ok = 0
    for _ in range(1500):
        vals = [repr(rng.choice([rng.random(), 0.1 + 0.2, rng.uniform(-1e6, 1e6), rng.random() * 1e18])) for _ in range(rng.randrange(1, 5))]
        rec = _with_values("[ " + " , ".join(vals) + " ]")           # whitespace: not canonical, same values
        st, out = canon(rec)
        want = port.encode(port.decode(rec))
        assert st in (0, 4), st
        if st == 0:
            assert out == want, (vals, out[:200])
            n_ok += 1
    assert n_ok > 1000
