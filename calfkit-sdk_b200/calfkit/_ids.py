"""Identifier sources used by the wire format.

The reference takes `uuid_utils.uuid7().hex` for CallFrame.frame_id, correlation ids and client
ids (reference calfkit/models/session_context.py:38, calfkit/client/base.py:83,
calfkit/client/client.py:103-104) and `pyd_ai_<uuid4hex>` for generated tool-call ids
(reference calfkit/_vendor/pydantic_ai/_utils.py:297-302).  uuid_utils is not in this image, so
uuid7 (RFC 9562 layout: 48-bit unix-ms | ver 7 | 12 rand | var 10 | 62 rand) is built here.

`set_id_source()` lets parity tests inject a deterministic generator, the same hook the oracle
installs into the reference (oracle/ref_harness.py:set_uuid_source).
"""
from __future__ import annotations

import os
import time
import uuid
from typing import Callable, Optional

_source: Optional[Callable[[], str]] = None


def set_id_source(fn: Optional[Callable[[], str]]) -> None:
    global _source
    _source = fn


def uuid7_hex() -> str:
    if _source is not None:
        return _source()
    ms = time.time_ns() // 1_000_000
    rnd = int.from_bytes(os.urandom(10), "big")
    rand_a = (rnd >> 62) & 0xFFF
    rand_b = rnd & ((1 << 62) - 1)
    v = ((ms & ((1 << 48) - 1)) << 80) | (0x7 << 76) | (rand_a << 64) | (0b10 << 62) | rand_b
    return f"{v:032x}"


def generate_tool_call_id() -> str:
    return f"pyd_ai_{uuid.uuid4().hex}"
