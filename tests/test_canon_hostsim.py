"""CPU fuzz of the device canonicaliser source (csrc/ck_canon.cuh, g++ build): its verdicts against the
reference codec (pydantic on the mirrored models, themselves pinned to the reference by the goldens).
  CK_OK (0)             -> bytes identical to dump(validate(input)), and the fast walker accepts them
  CK_JSON_INVALID (2)   -> pydantic raises exactly one json_invalid error
  CK_SCHEMA_INVALID (3) -> pydantic raises, and not for JSON syntax
  CK_UNSUPPORTED (4)    -> no claim (counted, must stay a minority on realistic inputs)"""
import json
import random

from conftest import as_bytes, golden
from hostsim import canon, walk, walk_trust
from pydantic import ValidationError


def _truth(b: bytes):
    from calfkit import _ids
    from calfkit.models import Envelope
    _ids.set_id_source(lambda: "0" * 31 + "1")
    try:
        return 0, Envelope.model_validate_json(b).model_dump_json().encode()
    except ValidationError as e:
        errs = e.errors()
        return (2 if errs[0]["type"] == "json_invalid" and len(errs) == 1 else 3), b""
    finally:
        _ids.set_id_source(None)


def _check(b: bytes, stats: dict):
    st, out = canon(b)
    stats[st] = stats.get(st, 0) + 1
    if st == 4:
        return
    tst, tout = _truth(b)
    assert st == tst, (st, tst, b[:400])
    if st == 0:
        assert out == tout, (b[:400], out[:200], tout[:200])
        assert walk_trust(out)[0], out[:400]    # whatever the canonicaliser emits, the second walk must accept


def test_goldens():
    stats: dict = {}
    for c in golden("codec.json") + golden("tool_node.json"):
        _check(as_bytes(c["input"]), stats)
    assert stats[0] > 100 and stats.get(4, 0) < 15


def test_reformatted_records_are_recovered():
    """pretty-printed, key-sorted and ASCII-escaped spellings of valid records canonicalise to the
    reference bytes (none of them is UNSUPPORTED)"""
    from calfkit import synth
    recs = synth.tool_events(20, seed=31) + synth.tool_events(10, seed=32, size=None, full_history=True) + \
        synth.fanout_events(2, seed=33, fanout=5) + synth.mixed_events(10, seed=34, hi=8000)
    for r in recs:
        obj = json.loads(r)
        for variant in (json.dumps(obj, indent=2), json.dumps(obj, sort_keys=True),
                        json.dumps(obj, ensure_ascii=True, separators=(" , ", " : "))):
            st, out = canon(variant.encode())
            tst, tout = _truth(variant.encode())       # (sort_keys also reorders Any dicts, whose order is preserved)
            assert st == 0 and tst == 0 and out == tout
            assert walk(out)[0]


def test_mutation_fuzz():
    from calfkit import synth
    rng = random.Random(3)
    seeds = [as_bytes(c["input"]) for c in golden("codec.json") + golden("tool_node.json") if 0 < len(as_bytes(c["input"])) < 5000]
    seeds += synth.tool_events(6, seed=9) + synth.tool_events(4, seed=8, size=None, full_history=True)
    for s in list(seeds):
        try:
            obj = json.loads(s)
        except Exception:
            continue
        seeds.append(json.dumps(obj, indent=1).encode())
        seeds.append(json.dumps(obj, sort_keys=True, ensure_ascii=True, separators=(", ", " : ")).encode())
    tok = [b'"', b"{", b"}", b"[", b"]", b",", b":", b"\\", b" ", b"\n", b"0", b"1", b"9", b"e", b"E", b".", b"-", b"+", b"null", b"true",
           b"false", b"1.5", b"1e5", b"-0", b"0.10", b'"a"', b"{}", b"[]", b"\\u0041", b"\\/", b"\xc3\xa9", b"\xff", b"\x01", b"\t",
           b'"kind":"tool-return",', b'"a":1,', b"NaN", b"Infinity", b"Z", b"+00:00", b".000000", b".5", b"00", b'"zz":[1,{"q":2}],',
           b'"part_kind":"text",', b'"kind":"request",', b"\\ud83d\\ude00", b"\\ud800", b"2.50", b"1E-7", b"12e3"]
    stats: dict = {}
    for _ in range(60000):
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
        if b:
            _check(bytes(b), stats)
    assert stats.get(0, 0) > 3000 and stats.get(3, 0) > 1000 and stats.get(2, 0) > 10000


def test_datetime_spellings():
    """separators T/t/space/_, fraction 1-9 digits after '.' or ',', Z/z, +HH:MM / +HHMM offsets: re-emitted exactly as pydantic
    does, everything else declared unsupported or invalid, never re-spelled differently"""
    from calfkit import synth
    r = synth.tool_events(1, seed=3)[0]
    i = r.index(b'"timestamp":"') + len(b'"timestamp":"')
    j = r.index(b'"', i)
    rng = random.Random(4)
    stats: dict = {}
    for _ in range(4000):
        sep = rng.choice(["T", "t", " ", "_", "T", "x"])
        frac = rng.choice(["", "", ".5", ",25", ".123456", ".1234567", ".000000", ".000", ".9999999999"])
        zone = rng.choice(["", "Z", "z", "+00:00", "-00:00", "+0000", "+05:30", "-0530", "+23:59", "+24:00", "+01", "+1:00", "Z "])
        s = f"{rng.randrange(1, 9999):04d}-{rng.randrange(0, 14):02d}-{rng.randrange(0, 33):02d}{sep}{rng.randrange(0, 25):02d}:{rng.randrange(0, 61):02d}:{rng.randrange(0, 61):02d}{frac}{zone}"
        _check(r[:i] + s.encode() + r[j:], stats)
    assert stats.get(0, 0) > 500
    assert stats.get(3, 0) > 500        # out-of-range fields in the RFC 3339 layout are the reference's parsing errors, not "unsupported"


def test_unix_timestamps_and_numeric_strings():
    """datetime fields given as unix numbers (seconds; milliseconds beyond the 2e10 watershed), as JSON numbers or as strings
    holding a plain integer: UTC, exactly as the reference dumps them; fractions / exponents stay declared-unsupported"""
    from calfkit import synth
    r = synth.tool_events(1, seed=3)[0]
    i = r.index(b'"timestamp":"') + len(b'"timestamp":')
    j = r.index(b'"', i + 1) + 1
    rng = random.Random(9)
    stats: dict = {}
    specials = [0, -1, 1, 86399, 86400, -86400, 951782400, 951868800, 1767225600, 20000000000, 20000000001, -20000000000, -20000000001,
                253402300799, 253402300800, -62135596800, -62135596801, 1767225600123, -1767225600123, 99999999999999, 4102444800, 68169600]
    for k in range(3000):
        v = rng.choice(specials) if k % 4 == 0 else rng.choice([rng.randrange(-3 * 10**10, 3 * 10**10), rng.randrange(-10**14, 10**14),
                                                              rng.randrange(0, 2 * 10**9), rng.randrange(-10**16, 10**16)])
        form = rng.randrange(6)
        lit = [str(v), str(v), '"' + str(v) + '"', str(v) + ".0", str(v) + ".5", "%de0" % v][form]
        _check(r[:i] + lit.encode() + r[j:], stats)
    assert stats.get(0, 0) > 1200 and stats.get(4, 0) > 300


def test_long_float_literals_are_respelled_as_the_reference_does():
    """float literals with 16-19 significant digits that are not the shortest round-trip spelling of their double (a value
    printed with %.17g or copied from a decimal source) come back as exactly what the reference dumps — csrc/ck_float.cuh
    finds that spelling by an exact search, no binary floating point involved; the re-emitted record is a fixed point"""
    import struct
    from calfkit import synth
    rng = random.Random(17)
    base = json.loads(synth.tool_events(1, seed=61)[0])
    lits = ["123456789.123456789", "0.1000000000000000055", "5.6843418860808015e-14", "0.30000000000000004", "1.0000000000000002",
            "123456789012345680.0", "9007199254740993.0", "2.2250738585072014e-280", "0.3000000000000000166", "4.35", "4.3499999999999996447",
            "1234567890123456789e-5", "72057594037927936.0", "9.999999999999999e22", "1.00000000000000011102230246251565404",
            # the edges: overflow is null, below half the smallest subnormal is zero, subnormals have their own shortest spelling
            "1e400", "-1e400", "1.7976931348623157e308", "1.7976931348623158e308", "1.797693134862315807e308", "1.797693134862315808e308",
            "5e-324", "4.9406564584124654e-324", "3e-324", "2.4703282292062327e-324", "2.4703282292062328e-324", "1e-400", "-1e-400",
            "2.2250738585072014e-308", "2.225073858507201e-308", "2.2250738585072011e-308", "1e-323", "9.8813129168249309e-324", "123e-320"]
    for _ in range(300):
        d = struct.unpack("<d", struct.pack("<Q", rng.getrandbits(64)))[0]
        if d != d or abs(d) == float("inf") or abs(d) < 1e-280 or abs(d) > 1e280:
            continue
        lits.append(("%." + str(rng.choice([16, 17, 18])) + "e") % d)
    decided = 0
    for k in range(0, len(lits), 8):
        chunk = lits[k:k + 8]
        txt = json.dumps(base).replace('"metadata": null', '"metadata": {"f": [' + ", ".join(chunk) + ']}', 1)
        assert '"f": [' in txt
        st, out = canon(txt.encode())
        tst, tout = _truth(txt.encode())
        assert tst == 0
        if st == 4:
            assert any(len(c.split("e")[0].replace(".", "").replace("-", "").strip("0")) > 19 for c in chunk), chunk   # only > 19 digits may be declined
            continue
        assert st == 0 and out == tout, (chunk, out[out.find(b'"f"'):][:300], tout[tout.find(b'"f"'):][:300])
        assert walk_trust(out)[0]
        decided += 1
    assert decided >= 35


def test_int_fields_given_as_other_number_spellings():
    """usage counters given as floats / strings: 7.0 is 7, 1.5 is the reference's int_from_float error (decided for <= 15
    significant digits without an exponent), the rest is declared unsupported — never a different integer"""
    from calfkit import synth
    r = synth.tool_events(1, seed=5, size=None, full_history=True)[0]
    k0 = r.index(b'"input_tokens":') + len(b'"input_tokens":')
    k1 = k0
    while r[k1:k1 + 1].isdigit():
        k1 += 1
    rng = random.Random(21)
    stats: dict = {}
    for _ in range(3000):
        ip = str(rng.choice([0, 1, 7, 51, 12345, 10**14, 10**15 + 3, rng.randrange(0, 10**rng.randrange(1, 17))]))
        fr = rng.choice(["", ".0", ".000", ".5", ".25", ".000001", ".10", "." + str(rng.randrange(1, 10**rng.randrange(1, 12))), ".0e0", "e2", ".5e1"])
        sign = rng.choice(["", "", "-"])
        lit = sign + ip + fr
        if rng.random() < 0.15:
            lit = '"' + lit + '"'
        _check(r[:k0] + lit.encode() + r[k1:], stats)
    assert stats.get(0, 0) > 500 and stats.get(3, 0) > 500


def test_shadowed_duplicate_dict_values_are_still_validated():
    """dict[str, Model]: the reference validates every member before the assignment, so a duplicate key whose EARLIER value is
    not a valid instance fails the record although the later value replaces it (found by scripts/fuzz_canon.py); for model
    fields the last one simply wins"""
    from calfkit import synth
    r = synth.tool_events(1, seed=5, size=None, full_history=True)[0].decode()
    k = r.index('"tool_calls":{') + len('"tool_calls":{')
    key = r[k:r.index('":{', k) + 1]
    end = r.index('},"tool_results"', k)
    stats: dict = {}
    for v in (r[:k] + key + ':{"art_kind":"tool-call"},' + r[k:],                 # invalid value shadowed by the valid one
              r[:end] + "," + key + ':{"art_kind":"tool-call"}' + r[end:],         # valid value shadowed by the invalid one
              r[:k] + key + ":" + r[k + len(key) + 1:end] + "," + r[k:],            # the same valid value twice
              r.replace('"details":{}', '"details":{"a":"x","a":1}', 1),            # dict[str, int] with an invalid shadowed value
              r.replace('"details":{}', '"details":{"a":1,"a":2}', 1),
              r.replace('"input_tokens":51', '"input_tokens":"x","input_tokens":51', 1)):   # model field: last wins, no error
        _check(v.encode(), stats)
    assert stats == {3: 3, 0: 3}
