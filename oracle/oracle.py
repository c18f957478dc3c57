"""ctypes binding of the CPU oracle (oracle/sais_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  Nothing under suffix_b200/
imports this module.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")


def build(force: bool = False) -> str:
    """Compile liboracle.so with gcc if missing or stale."""
    srcs = [os.path.join(_HERE, f) for f in ("sais_oracle.c", "sais_level.inc", "Makefile")]
    stale = (not os.path.exists(_SO)) or any(
        os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        u8p = ctypes.c_void_p
        u32p = ctypes.c_void_p
        u64 = ctypes.c_uint64
        for name in ("oracle_sais", "oracle_naive_sa"):
            getattr(L, name).argtypes = [u8p, u64, u32p]
            getattr(L, name).restype = ctypes.c_int
        L.oracle_types.argtypes = [u8p, u64, u8p]
        L.oracle_types.restype = ctypes.c_int
        for name in ("oracle_lcp_quadratic", "oracle_lcp_lens", "oracle_lcp_kasai"):
            getattr(L, name).argtypes = [u8p, u64, u32p, u32p]
            getattr(L, name).restype = ctypes.c_int
        L.oracle_verify_sa.argtypes = [u8p, u64, u32p]
        L.oracle_verify_sa.restype = ctypes.c_int64
        L.oracle_positions.argtypes = [u8p, u64, u32p, u8p, u64,
                                       ctypes.POINTER(u64), ctypes.POINTER(u64)]
        L.oracle_positions.restype = ctypes.c_int
        L.oracle_any_position.argtypes = [u8p, u64, u32p, u8p, u64,
                                          ctypes.POINTER(ctypes.c_uint32)]
        L.oracle_any_position.restype = ctypes.c_int
        _lib = L
    return _lib


def _as_u8(text) -> np.ndarray:
    if isinstance(text, str):
        text = text.encode("utf-8")
    if isinstance(text, (bytes, bytearray, memoryview)):
        return np.frombuffer(bytes(text), dtype=np.uint8)
    a = np.ascontiguousarray(text, dtype=np.uint8)
    return a


def _ptr(a: np.ndarray):
    return ctypes.c_void_p(a.ctypes.data)


def sais(text) -> np.ndarray:
    """Restated reference sais_table (src/table.rs:378-386)."""
    t = _as_u8(text)
    sa = np.zeros(len(t), dtype=np.uint32)
    rc = lib().oracle_sais(_ptr(t), len(t), _ptr(sa))
    assert rc == 0
    return sa


def naive_sa(text) -> np.ndarray:
    """Restated reference naive_table (src/table.rs:367-376)."""
    t = _as_u8(text)
    sa = np.zeros(len(t), dtype=np.uint32)
    rc = lib().oracle_naive_sa(_ptr(t), len(t), _ptr(sa))
    assert rc == 0
    return sa


def types(text) -> np.ndarray:
    """0 Ascending(S) / 1 Descending(L) / 2 Valley(LMS) per byte (src/table.rs:592-615)."""
    t = _as_u8(text)
    out = np.zeros(len(t), dtype=np.uint8)
    rc = lib().oracle_types(_ptr(t), len(t), _ptr(out))
    assert rc == 0
    return out


def _lcp(fn, text, sa) -> np.ndarray:
    t = _as_u8(text)
    sa = np.ascontiguousarray(sa, dtype=np.uint32)
    assert len(sa) == len(t)
    out = np.zeros(len(t), dtype=np.uint32)
    rc = fn(_ptr(t), len(t), _ptr(sa), _ptr(out))
    assert rc == 0
    return out


def lcp_quadratic(text, sa) -> np.ndarray:
    """src/table.rs:348-361."""
    return _lcp(lib().oracle_lcp_quadratic, text, sa)


def lcp_lens(text, sa) -> np.ndarray:
    """src/table.rs:130-138 (with the wasted inverse fill; for timing)."""
    return _lcp(lib().oracle_lcp_lens, text, sa)


def lcp_kasai(text, sa) -> np.ndarray:
    return _lcp(lib().oracle_lcp_kasai, text, sa)


def verify_sa(text, sa) -> int:
    """O(n) check that `sa` is THE suffix array of `text` (permutation + neighbour
    order through the inverse); 0 = valid, else 1 + index of the first violation."""
    t = _as_u8(text)
    sa = np.ascontiguousarray(sa, dtype=np.uint32)
    assert len(sa) == len(t)
    rc = int(lib().oracle_verify_sa(_ptr(t), len(t), _ptr(sa)))
    assert rc >= 0, "oracle_verify_sa: out of memory"
    return rc


def positions(text, sa, query):
    """src/table.rs:223-259 -> (start, end) range into sa."""
    t = _as_u8(text)
    q = _as_u8(query)
    sa = np.ascontiguousarray(sa, dtype=np.uint32)
    s = ctypes.c_uint64(0)
    e = ctypes.c_uint64(0)
    rc = lib().oracle_positions(_ptr(t), len(t), _ptr(sa), _ptr(q), len(q),
                                ctypes.byref(s), ctypes.byref(e))
    assert rc == 0
    return int(s.value), int(e.value)


def any_position(text, sa, query):
    """src/table.rs:279-293 -> position or None."""
    t = _as_u8(text)
    q = _as_u8(query)
    sa = np.ascontiguousarray(sa, dtype=np.uint32)
    pos = ctypes.c_uint32(0)
    hit = lib().oracle_any_position(_ptr(t), len(t), _ptr(sa), _ptr(q), len(q),
                                    ctypes.byref(pos))
    return int(pos.value) if hit else None
