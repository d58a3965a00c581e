"""CPU: the C-ABI library loads and exports every symbol include/ecb200.h declares
(no compute calls: there is no GPU here), and argument validation that does not need a
device behaves."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    hdr = open(os.path.join(ROOT, "include", "ecb200.h")).read()
    return sorted(set(re.findall(r"^\s*(?:int|uint64_t|const char\*)\s+(ecb200_\w+)\s*\(", hdr, re.M)))


def test_header_declares_expected_entry_points():
    names = _declared()
    for must in ("ecb200_scalar_mult_p256", "ecb200_mgry_mul", "ecb200_mgry_sqr", "ecb200_mgry_add", "ecb200_mgry_sub",
                 "ecb200_zdau", "ecb200_zaddu", "ecb200_dblu", "ecb200_add_z2_1", "ecb200_trplu", "ecb200_to_affine"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from ecsimd_b200 import capi
    lib = capi.load()            # raises loudly if the .so is missing
    for name in _declared():
        assert hasattr(lib, name), name
        assert name in capi.SYMBOLS, "binding missing for %s" % name
    assert lib.ecb200_abi_version() == 1


def test_argument_validation_without_device():
    from ecsimd_b200 import capi
    lib = capi.load()
    # unknown layout / PACK4 with n % 4 != 0 are rejected before any CUDA call
    buf = (C.c_uint32 * 64)()
    assert lib.ecb200_mgry_mul(buf, buf, buf, 4, 7, None) == -1
    assert b"layout" in lib.ecb200_last_error()
    assert lib.ecb200_mgry_mul(buf, buf, buf, 3, capi.LAYOUT_PACK4, None) == -1
    assert lib.ecb200_mgry_shift_left(buf, buf, 0, 4, 0, None) == -1
    assert lib.ecb200_mgry_mul(buf, buf, buf, 0, 0, None) == 0          # empty batch is a no-op


def test_product_does_not_link_the_oracle():
    """the shipped library must not depend on, or contain, the CPU checker"""
    import subprocess
    from ecsimd_b200 import capi
    out = subprocess.run(["nm", "-D", capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "orc_" not in out and "ref_" not in out
    for mod in ("capi.py", "host.py", "device.py", "__init__.py", "shard.py"):
        p = os.path.join(ROOT, "ecsimd_b200", mod)
        if os.path.exists(p):
            src = open(p).read()
            assert "oracle" not in src.replace("no CPU fallback", ""), mod
