"""Fuzz of ckf_shortest / ckf_nearest_double_ex (host build of csrc/ck_float.cuh) against repr(float): subnormals, the smallest\nnormals, the largest doubles, overflow and underflow, literals of 1..19 digits with perturbed tails.  usage: fuzz_floats.py [seed] [n]"""
import sys, random, struct, ctypes
import os; sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import hostsim
hostsim.build(force=True)
L = hostsim.lib()
L.ck_host_shortest.argtypes = [ctypes.c_uint64, ctypes.c_int, ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_int)]
L.ck_host_shortest.restype = ctypes.c_int
from decimal import Decimal, getcontext
getcontext().prec = 60
rng = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
def repr_digits(d):
    t = Decimal(repr(d)).as_tuple()
    digs = "".join(map(str, t.digits)).lstrip("0") or "0"
    exp = t.exponent
    while len(digs) > 1 and digs.endswith("0"): digs = digs[:-1]; exp += 1
    return int(digs), exp
n_ok = n_und = 0
N = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
for it in range(N):
    mode = rng.randrange(6)
    if mode == 0: d = struct.unpack("<d", struct.pack("<Q", rng.getrandbits(52)))[0]                    # subnormals
    elif mode == 1: d = struct.unpack("<d", struct.pack("<Q", rng.getrandbits(52) | (rng.randrange(1, 4) << 52)))[0]   # smallest normals
    elif mode == 2: d = struct.unpack("<d", struct.pack("<Q", rng.getrandbits(52) | (rng.randrange(2043, 2047) << 52)))[0]  # largest
    elif mode == 3: d = struct.unpack("<d", struct.pack("<Q", rng.getrandbits(63)))[0]
    elif mode == 4: d = rng.choice([5e-324, 1e-323, 2.2250738585072014e-308, 2.225073858507201e-308, 1.7976931348623157e308, 4.9406564584124654e-324 * rng.randrange(1, 1000)])
    else: d = float(rng.randrange(1, 10**rng.randrange(1, 16))) * 10.0 ** rng.randrange(-330, 300)
    if d != d or d == float("inf") or d == 0.0: continue
    nd = rng.choice([1, 2, 3, 5, 8, 12, 15, 16, 17, 17, 18, 19])
    lit = ("%." + str(nd - 1) + "e") % d
    mant, ex = lit.split("e")
    digs = mant.replace(".", "")
    if rng.random() < 0.3 and nd > 1:
        k = rng.randrange(max(1, nd - 3), nd)
        digs = digs[:k] + "".join(rng.choice("0123456789") for _ in range(nd - k))
    m = int(digs); k = int(ex) - (nd - 1)
    if m == 0: continue
    while m % 10 == 0: m //= 10; k += 1
    try:
        dd = float(Decimal(m).scaleb(k))
    except OverflowError:
        dd = float("inf")
    ms, ks = ctypes.c_uint64(0), ctypes.c_int(0)
    ok = L.ck_host_shortest(m, k, ctypes.byref(ms), ctypes.byref(ks))
    if dd == float("inf") or dd == 0.0:
        if ok: print("WRONG: decided an inf/zero", m, k); sys.exit(1)
        continue
    want = repr_digits(dd)
    if ok:
        n_ok += 1
        if (ms.value, ks.value) != want:
            print("WRONG", m, k, dd, (ms.value, ks.value), want); sys.exit(1)
    else:
        n_und += 1
        if n_und < 5: print("undecided", m, k, dd)
print("ok: decided", n_ok, "undecided", n_und)
