"""GPU: the C++ drop-in header (include/ecsimd_b200/ecsimd.hpp) runs the reference's own
P-256 known-answer tests (tests/curve_group.cpp, tests/curve_point.cpp of the reference)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build():
    exe = os.path.join(ROOT, "build", "test_shim")
    src = os.path.join(ROOT, "tests", "cpp", "test_shim.cpp")
    lib = os.path.join(ROOT, "ecsimd_b200")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.run(["g++", "-std=c++17", "-O1", "-o", exe, src, "-L" + lib, "-lecb200", "-Wl,-rpath," + lib], check=True)
    return exe


def test_cpp_shim_compiles():
    """not gpu: the header is self-contained C++17 and links against the C ABI"""
    _build()


@pytest.mark.gpu
def test_cpp_shim_reference_kats():
    exe = _build()
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip().startswith("ok"), out.stdout + out.stderr
