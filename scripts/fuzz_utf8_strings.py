"""Targeted fuzz of the walker's string core (host build of csrc/ck_walk.cuh): valid and invalid UTF-8 sequences and escapes at
every alignment inside the 8-byte scan words; the walker must accept exactly the fixed points of the reference codec."""
import os, sys, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p_ in ("tests", "calfkit-sdk_b200", ""): sys.path.insert(0, os.path.join(ROOT, p_))
from hostsim import walk
from calfkit import synth
from pydantic import ValidationError
from calfkit.models import Envelope
def is_fixed(b):
    try: return Envelope.model_validate_json(b).model_dump_json().encode() == b
    except ValidationError: return False
rng = random.Random(11)
base = synth.tool_events(4, seed=3)[0]
assert walk(base)[0]
# a string content spot: the location argument value
k = base.index(b'"location":"') + len(b'"location":"')
e = base.index(b'"', k)
pieces_valid = ["é", "ü", "—", "漢", "字", "🙂", "a", "bc", "\\n", "\\t", "\\\"", "\\\\", "ß", "€", "\U0010ffff", "ࠀ", "퟿", "", "x" * 7]
bad_bytes = [b"\xc0\x80", b"\xc1\xbf", b"\xe0\x80\x80", b"\xe0\x9f\xbf", b"\xed\xa0\x80", b"\xed\xbf\xbf", b"\xf0\x80\x80\x80", b"\xf0\x8f\xbf\xbf",
             b"\xf4\x90\x80\x80", b"\xf5\x80\x80\x80", b"\x80", b"\xbf", b"\xc3", b"\xe2\x82", b"\xf0\x9f\x99", b"\xc3\x28", b"\xe2\x28\xa1", b"\xe2\x82\x28",
             b"\xf0\x28\x8c\xbc", b"\xf0\x90\x28\xbc", b"\xf0\x28\x8c\x28", b"\xff", b"\xfe", b"\x1f", b"\\x", b"\\u0041", b"\\u001f", b"\\u000a", b"\\"]
n_acc = n_rej = 0
for it in range(60000):
    parts = []
    for _ in range(rng.randrange(1, 12)):
        if rng.random() < 0.12:
            parts.append(rng.choice(bad_bytes))
        else:
            parts.append(rng.choice(pieces_valid).encode())
    pad = b"a" * rng.randrange(0, 9)               # every alignment of the specials within the 8-byte words
    content = pad + b"".join(parts)
    rec = base[:k] + content + base[e:]
    acc = walk(rec)[0]
    fx = is_fixed(rec)
    if acc and not fx:
        print("UNSOUND", rec[k - 5:k + len(content) + 5]); sys.exit(1)
    if fx and not acc:
        print("INCOMPLETE", rec[k - 5:k + len(content) + 5]); sys.exit(1)
    n_acc += acc; n_rej += (not acc)
print("ok", n_acc, "accepted", n_rej, "rejected; walker == fixed-point test on all")
