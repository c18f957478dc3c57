"""CPU-only checks of the boundary: the library loads, exports every symbol
that include/*.h declares, and fails loudly (no fallback) without a GPU."""
import ctypes
import os
import re

import pytest

from suffix_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    names = set()
    for h in os.listdir(os.path.join(ROOT, "include")):
        if not h.endswith(".h"):
            continue
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(b200sa_[a-z0-9_]+)\s*\(", src))
    return sorted(names)


def test_exports_every_declared_symbol():
    L = ctypes.CDLL(_lib.LIB_PATH)
    decl = _declared()
    assert len(decl) >= 15
    for name in decl:
        assert hasattr(L, name), name


def test_strerror_and_version():
    L = _lib.lib()
    assert L.b200sa_strerror(0) == b"ok"
    assert b"2^32" in L.b200sa_strerror(-2)
    assert b"sm_100a" in L.b200sa_version()


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.B200SAError) as e:
        _lib.Context(0)
    assert e.value.code == -3
    from suffix_b200 import SuffixTable
    with pytest.raises(_lib.B200SAError):
        SuffixTable("banana")


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "suffix_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.lower() or f == "gen.py", (f, "product code must not reference oracle/")


def test_header_documents_max_n():
    src = open(os.path.join(ROOT, "include", "b200sa.h")).read()
    assert "B200SA_MAX_N 0xFFFFF000ull" in src
