"""TEST INFRASTRUCTURE ONLY: g++ build of the device walker for CPU-side fuzzing."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(HERE, "libck_hostsim.so")


def build(force: bool = False) -> str:
    srcs = [os.path.join(HERE, "hostsim.cpp"),
            os.path.join(HERE, "..", "..", "calfkit-sdk_b200", "csrc", "ck_walk.cuh"),
            os.path.join(HERE, "..", "..", "calfkit-sdk_b200", "csrc", "ck_common.h"),
            os.path.join(HERE, "ck_vm.cuh"),
            os.path.join(HERE, "ck_vm_prog.h"),
            os.path.join(HERE, "..", "..", "calfkit-sdk_b200", "csrc", "ck_canon.cuh"),
            os.path.join(HERE, "..", "..", "calfkit-sdk_b200", "csrc", "ck_float.cuh")]
    if force or not os.path.exists(_LIB) or any(os.path.getmtime(s) > os.path.getmtime(_LIB) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wno-unknown-pragmas",
                               "-I", os.path.join(HERE, "..", "..", "calfkit-sdk_b200", "csrc"), "-o", _LIB, srcs[0]])
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.ck_host_walk.argtypes = [ctypes.c_char_p, ctypes.c_uint32, ctypes.c_void_p]
        _lib.ck_host_walk.restype = ctypes.c_int
        _lib.ck_host_walk_global.argtypes = [ctypes.c_char_p, ctypes.c_uint32, ctypes.c_void_p]
        _lib.ck_host_walk_global.restype = ctypes.c_int
        _lib.ck_host_walk_trust.argtypes = [ctypes.c_char_p, ctypes.c_uint32, ctypes.c_void_p]
        _lib.ck_host_walk_trust.restype = ctypes.c_int
        _lib.ck_host_walk_skip.argtypes = [ctypes.c_char_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32]
        _lib.ck_host_walk_skip.restype = ctypes.c_int
        _lib.ck_host_elem.argtypes = [ctypes.c_char_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32]
        _lib.ck_host_elem.restype = ctypes.c_int
        _lib.ck_host_num_cols.restype = ctypes.c_int
        _lib.ck_host_vm_walk.argtypes = [ctypes.c_char_p, ctypes.c_uint32, ctypes.c_void_p]
        _lib.ck_host_vm_walk.restype = ctypes.c_int
        _lib.ck_host_canon.argtypes = [ctypes.c_char_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p]
        _lib.ck_host_canon.restype = ctypes.c_int
    return _lib


def walk(payload: bytes):
    """-> (accepted: bool, cols: np.ndarray[uint32])"""
    L = lib()
    cols = np.zeros(L.ck_host_num_cols(), dtype=np.uint32)
    ok = L.ck_host_walk(payload, len(payload), cols.ctypes.data)
    return bool(ok), cols


def walk_skip(payload: bytes, open_: int, close: int):
    """the record walk with message_history = [open_, close] handed to the element pass: -> (accepted, cols)"""
    L = lib()
    cols = np.zeros(L.ck_host_num_cols(), dtype=np.uint32)
    ok = L.ck_host_walk_skip(payload, len(payload), cols.ctypes.data, open_, close)
    return bool(ok), cols


def elem_ok(payload: bytes, start: int, end: int) -> bool:
    """what ck_walk_elems_kernel decides for one listed message"""
    return bool(lib().ck_host_elem(payload, len(payload), start, end))


def walk_global(payload: bytes):
    """the same walker over the global-load reader (GRd): -> (accepted, cols)"""
    L = lib()
    cols = np.zeros(L.ck_host_num_cols(), dtype=np.uint32)
    ok = L.ck_host_walk_global(payload, len(payload), cols.ctypes.data)
    return bool(ok), cols


def walk_trust(payload: bytes):
    """the second walk of the decode pass (WRdT): -> (accepted, cols)"""
    L = lib()
    cols = np.zeros(L.ck_host_num_cols(), dtype=np.uint32)
    ok = L.ck_host_walk_trust(payload, len(payload), cols.ctypes.data)
    return bool(ok), cols


def decode(payload: bytes):
    """the device decode pass as ck_submit runs it: walk; if not proven canonical, canonicalise and walk the re-emitted
    bytes with the trusting reader.  -> (status 0 ok / other, canonical bytes or b"")"""
    L = lib()
    ok, _ = walk(payload)
    if ok:
        return 0, payload
    st, out = canon(payload)
    if st != 0:
        return st, b""
    cols = np.zeros(L.ck_host_num_cols(), dtype=np.uint32)
    return (0 if L.ck_host_walk_trust(out, len(out), cols.ctypes.data) else 4), out


def vm_walk(payload: bytes):
    """the table-driven walker (csrc/ck_vm.cuh): -> (accepted, cols)"""
    L = lib()
    cols = np.zeros(L.ck_host_num_cols(), dtype=np.uint32)
    ok = L.ck_host_vm_walk(payload, len(payload), cols.ctypes.data)
    return bool(ok), cols


def canon(payload: bytes):
    """the device canonicaliser (csrc/ck_canon.cuh): -> (status, canonical bytes or b"")"""
    L = lib()
    cap = 4 * len(payload) + 8192
    out = np.zeros(cap, dtype=np.uint8)
    n = ctypes.c_uint32(0)
    st = L.ck_host_canon(payload, len(payload), out.ctypes.data, cap, ctypes.byref(n))
    return st, out[:n.value].tobytes()
