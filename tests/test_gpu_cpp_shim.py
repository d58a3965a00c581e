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


def _prebuilt(name):
    exe = os.path.join(ROOT, "oracle", "_ref", name)
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/%s was not built (it needs the reference checkout at build time)" % name)
    return exe


@pytest.mark.gpu
def test_reference_test_tus_unmodified_on_the_engine():
    """the reference's OWN tests/curve_group.cpp and tests/curve_point.cpp, compiled unmodified against the mirror
    header (oracle/Makefile: _ref/ref_tests_on_b200): DBLU, ZADDU, ZDAU, Swap, ScalarMult, FromX, ToFromAffine"""
    out = subprocess.run([_prebuilt("ref_tests_on_b200")], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "ok 7 tests" in out.stdout, out.stdout + out.stderr


@pytest.mark.gpu
def test_integration_binding_against_the_real_reference_headers():
    """INTEGRATION.md section 1 as a real translation unit (oracle/integration_b200.cpp): reference types in, GPU results
    bit-identical to the reference's own CPU code on the KAT scalars and a seeded batch of packs"""
    out = subprocess.run([_prebuilt("integration_b200")], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and out.stdout.strip().splitlines()[-1].startswith("ok"), out.stdout + out.stderr


@pytest.mark.gpu
def test_cpp_shim_reference_kats():
    exe = _build()
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip().startswith("ok"), out.stdout + out.stderr
