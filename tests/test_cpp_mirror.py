"""C++ host mirror (include/b200sa_table.hpp): compiles on CPU; the KATs of
tests/tests.rs run through it on the GPU box."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "test_table")


def _build():
    src = os.path.join(ROOT, "tests", "cpp", "test_table.cpp")
    cmd = ["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), src, "-o", EXE,
           "-L", os.path.join(ROOT, "suffix_b200"), "-lb200sa", "-Wl,-rpath," + os.path.join(ROOT, "suffix_b200")]
    subprocess.check_call(cmd)


def test_cpp_mirror_compiles():
    _build()
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_cpp_mirror_kats():
    _build()
    out = subprocess.run([EXE], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "cpp mirror ok" in out.stdout, out.stdout + out.stderr
