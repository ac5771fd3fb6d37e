"""The long-record decode (csrc/ck_walk_long.cuh) restated on the host build of the same walker sources: a record is accepted
when the record walk — which jumps over message_history from a PROPOSED '[' to a PROPOSED ']' — accepts and every proposed
message [start, end) is exactly one canonical ModelMessage.  The proposals come from a data-parallel pre-scan on the device;
here they come from a naive scanner that is deliberately easy to fool, and from an adversary.  Soundness must not depend on
them: whatever is accepted is a fixed point of the reference codec."""
import random

from hostsim import elem_ok, walk, walk_skip
from pydantic import ValidationError

MARK = b'"message_history":['


def _is_fixed(b: bytes) -> bool:
    from calfkit.models import Envelope
    try:
        return Envelope.model_validate_json(b).model_dump_json().encode() == b
    except ValidationError:
        return False


def _propose(b: bytes):
    """what the pre-scan proposes: the first '[' preceded by the key, its matching closer by bracket counting outside
    strings (backslash handling as crude as the device's), the commas at that level"""
    k = b.find(MARK)
    if k < 0:
        return None
    open_ = k + len(MARK) - 1
    depth, in_str, i, seps = 0, False, open_, []
    while i < len(b):
        c = b[i]
        if in_str:
            if c == 0x5C:
                i += 1
            elif c == 0x22:
                in_str = False
        elif c == 0x22:
            in_str = True
        elif c in b"[{":
            depth += 1
        elif c in b"]}":
            depth -= 1
            if depth == 0:
                return open_, i, seps
        elif c == 0x2C and depth == 1:
            seps.append(i)
        i += 1
    return None


def _pipeline_accepts(b: bytes, prop) -> bool:
    if prop is None:
        return walk(b)[0]
    open_, close, seps = prop
    if close <= open_ + 1:
        return walk(b)[0]
    ok, _ = walk_skip(b, open_, close)
    if not ok:
        return False
    starts = [open_ + 1] + [s + 1 for s in seps]
    ends = seps + [close]
    return all(elem_ok(b, s, e) for s, e in zip(starts, ends))


def _long_records():
    from calfkit import synth
    return [r for r in synth.mixed_events(60, seed=31, lo=4000, hi=30000, n_tools=1) if len(r) > 4000][:12]


def test_same_columns_with_and_without_the_jump():
    for r in _long_records():
        prop = _propose(r)
        assert prop is not None and len(prop[2]) >= 1
        ok0, c0 = walk(r)
        ok1, c1 = walk_skip(r, prop[0], prop[1])
        assert ok0 and ok1 and (c0 == c1).all()
        assert _pipeline_accepts(r, prop)
        # a jump offered at the wrong place is not taken (the walker only trusts a '[' it has reached itself)
        ok2, c2 = walk_skip(r, prop[0] + 1, prop[1])
        assert ok2 and (c2 == c0).all()
        # a jump to a place without the closer fails the record walk
        assert not walk_skip(r, prop[0], prop[1] - 1)[0]


def test_mutants_and_adversarial_proposals_are_sound():
    rng = random.Random(5)
    seeds = _long_records()
    tok = [b'"', b"{", b"}", b"[", b"]", b",", b":", b"\\", b" ", b"null", b'"a":1,', b"\xc3\xa9", b"\xff", b"\\\"", b"],[", b"},{",
           b'"part_kind":"text"', b'"kind":"request"', b"\\\\", b'\\\\"']
    accepted = fixed = 0
    for it in range(4000):
        b = bytearray(rng.choice(seeds))
        for _ in range(rng.choice([1, 1, 2, 3])):
            op, i = rng.randrange(5), rng.randrange(len(b))
            if op == 0:
                b[i] = rng.randrange(256)
            elif op == 1:
                del b[i]
            elif op == 2:
                b[i:i] = rng.choice(tok)
            elif op == 3:
                j = min(len(b), i + rng.randrange(1, 200)); del b[i:j]
            else:
                j = min(len(b), i + rng.randrange(1, 200)); b[i:i] = b[i:j]
        m = bytes(b)
        prop = _propose(m)
        if prop is not None and rng.random() < 0.3 and prop[2]:
            # the adversary: drop / shift separators, move the closer
            open_, close, seps = prop
            seps = list(seps)
            k = rng.randrange(3)
            if k == 0:
                del seps[rng.randrange(len(seps))]
            elif k == 1:
                j = rng.randrange(len(seps)); seps[j] = max(open_ + 1, min(close - 1, seps[j] + rng.choice([-3, -1, 1, 2, 7])))
                seps.sort()
            else:
                close = max(open_ + 2, min(len(m) - 1, close + rng.choice([-5, -1, 1, 4])))
                seps = [s for s in seps if s < close]
            prop = (open_, close, seps)
        if _pipeline_accepts(m, prop):
            accepted += 1
            assert _is_fixed(m), (it, m[:300])
        fixed += 1
    assert accepted > 20
