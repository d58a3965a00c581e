"""ctypes binding of the C ABI (include/ecb200.h -> ecsimd_b200/libecb200.so).

This is the same binding a maintainer of the reference would write against the
shared library; the Python layer adds no arithmetic.  There
is no CPU fallback: if the CUDA library is missing or a call fails, an exception
is raised.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ECB200_LIB", os.path.join(HERE, "libecb200.so"))  # override: A/B builds during development

LAYOUT_LANE, LAYOUT_PACK4, LAYOUT_SOA = 0, 1, 2
MEM_HOST, MEM_DEVICE = 0x00, 0x10
NO_QUIRK = 0x100
LAYOUTS = {"lane": LAYOUT_LANE, "pack4": LAYOUT_PACK4, "soa": LAYOUT_SOA}

# every symbol include/ecb200.h declares: (name, restype, argtypes)
_vp, _sz, _u32, _i = C.c_void_p, C.c_size_t, C.c_uint32, C.c_int
SYMBOLS = {
    "ecb200_abi_version": (_i, []),
    "ecb200_init": (_i, [_i]),
    "ecb200_init_devices": (_i, [C.POINTER(C.c_int), _i]),
    "ecb200_device_count": (_i, []),
    "ecb200_shutdown": (_i, []),
    "ecb200_last_error": (C.c_char_p, []),
    "ecb200_launch_count": (C.c_uint64, []),
    "ecb200_mgry_add": (_i, [_vp, _vp, _vp, _sz, _u32, _vp]),
    "ecb200_mgry_sub": (_i, [_vp, _vp, _vp, _sz, _u32, _vp]),
    "ecb200_mgry_mul": (_i, [_vp, _vp, _vp, _sz, _u32, _vp]),
    "ecb200_mgry_sqr": (_i, [_vp, _vp, _sz, _u32, _vp]),
    "ecb200_mgry_shift_left": (_i, [_vp, _vp, _i, _sz, _u32, _vp]),
    "ecb200_gfp_opposite": (_i, [_vp, _vp, _sz, _u32, _vp]),
    "ecb200_from_classical": (_i, [_vp, _vp, _sz, _u32, _vp]),
    "ecb200_to_classical": (_i, [_vp, _vp, _sz, _u32, _vp]),
    "ecb200_gfp_inverse": (_i, [_vp, _vp, _sz, _u32, _vp]),
    "ecb200_mgry_mul_chain": (_i, [_vp, _vp, _vp, _i, _sz, _u32, _vp]),
    "ecb200_dblu": (_i, [_vp, _vp, _vp, _sz, _u32, _vp]),
    "ecb200_zaddu": (_i, [_vp, _vp, _vp, _vp, _sz, _u32, _vp]),
    "ecb200_zdau": (_i, [_vp, _vp, _vp, _vp, _sz, _u32, _vp]),
    "ecb200_add_z2_1": (_i, [_vp, _vp, _vp, _sz, _u32, _vp]),
    "ecb200_trplu": (_i, [_vp, _vp, _vp, _sz, _u32, _vp]),
    "ecb200_scalar_mult_p256": (_i, [_vp, _vp, _vp, _sz, _u32, _vp]),
    "ecb200_scalar_mult_p256_base": (_i, [_vp, _vp, _sz, _u32, _vp]),
    "ecb200_scalar_mult_p256_1s": (_i, [_vp, _vp, _vp, _sz, _u32, _vp]),
    "ecb200_from_affine": (_i, [_vp, _vp, _sz, _u32, _vp]),
    "ecb200_to_affine": (_i, [_vp, _vp, _sz, _u32, _vp]),
    "ecb200_scalar_mult_p256_affine": (_i, [_vp, _vp, _vp, _sz, _u32, _vp]),
    "ecb200_gen_mod_add": (_i, [_vp, _vp, _vp, _vp, _sz, _u32, _vp]),
    "ecb200_gen_mod_sub": (_i, [_vp, _vp, _vp, _vp, _sz, _u32, _vp]),
    "ecb200_gen_mod_shift_left_one": (_i, [_vp, _vp, _vp, _sz, _u32, _vp]),
    "ecb200_gen_mgry_mul": (_i, [_vp, _vp, _vp, _vp, _sz, _u32, _vp]),
    "ecb200_gen_mgry_sqr": (_i, [_vp, _vp, _vp, _sz, _u32, _vp]),
    "ecb200_gen_from_classical": (_i, [_vp, _vp, _vp, _sz, _u32, _vp]),
    "ecb200_gen_to_classical": (_i, [_vp, _vp, _vp, _sz, _u32, _vp]),
    "ecb200_gen_mgry_pow": (_i, [_vp, _vp, _vp, _vp, _sz, _u32, _vp]),
    "ecb200_gen_opposite": (_i, [_vp, _vp, _vp, _sz, _u32, _vp]),
    "ecb200_mul512": (_i, [_vp, _vp, _vp, _sz, _u32, _vp]),
    "ecb200_square512": (_i, [_vp, _vp, _sz, _u32, _vp]),
    "ecb200_convert_layout": (_i, [_vp, _u32, _vp, _u32, _i, _sz, _u32, _vp]),
    "ecb200_bn_from_bytes_be": (_i, [_vp, _vp, _i, _sz, _u32, _vp]),
    "ecb200_bn_to_bytes_be": (_i, [_vp, _vp, _i, _sz, _u32, _vp]),
    "ecb200_from_x": (_i, [_vp, _vp, _vp, _sz, _u32, _vp]),
    "ecb200_synth_values": (_i, [_vp, C.c_uint64, C.c_uint64, _i, _sz, _u32, _vp]),
    "ecb200_checksum": (_i, [_vp, _vp, _sz, _vp]),
    "ecb200_microbench_mix": (_i, [_i, _i, _i, _i, C.POINTER(C.c_int), C.POINTER(C.c_float), _vp]),
    "ecb200_microbench_mix_count": (_i, []),
    "ecb200_microbench": (_i, [_i, _i, _i, _i, C.POINTER(C.c_double), C.POINTER(C.c_float), _vp]),
}


class Ecb200Error(RuntimeError):
    pass


_lib = None


def load():
    """dlopen the engine; fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise Ecb200Error(
            "%s is missing: build it with `python -m ecsimd_b200.build` (needs nvcc). "
            "There is no CPU fallback for this engine." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the export is missing
        fn.restype = res
        fn.argtypes = args
    if lib.ecb200_abi_version() != 1:
        raise Ecb200Error("ABI version mismatch")
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise Ecb200Error("ecb200 error %d: %s" % (rc, load().ecb200_last_error().decode()))


def init(device=0):
    check(load().ecb200_init(device))


def init_devices(devices):
    """single-process multi-GPU: host-memory batches of the scalar multiplication are cut over `devices`
    (include/ecb200.h: ecb200_init_devices); an empty list or one device restores single-device dispatch"""
    devices = list(devices)
    arr = (C.c_int * max(1, len(devices)))(*devices)
    check(load().ecb200_init_devices(arr, len(devices)))


def device_count():
    return int(load().ecb200_device_count())


def shutdown():
    """release the staging pool and the fixed-base tables of the current device (rebuilt on demand)"""
    check(load().ecb200_shutdown())


def launch_count():
    return int(load().ecb200_launch_count())


def _p(x):
    """pointer of a numpy array (host) or of anything with data_ptr() (torch tensor)"""
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        assert x.flags["C_CONTIGUOUS"]
        return x.ctypes.data
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    return int(x)


def call(name, *args):
    check(getattr(load(), name)(*args))
