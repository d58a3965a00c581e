"""ecsimd_b200 -- a B200-native batched NIST P-256 engine behind ecsimd's API.

The product is the CUDA shared library `libecb200.so` (C ABI in
include/ecb200.h); this package is the thin host-side mirror of the reference's
operator interface over it:

  * `ecsimd_b200.capi`   -- ctypes binding, one symbol per C-ABI entry point;
  * `ecsimd_b200.host`   -- numpy (host memory) front-end with the reference's
                            names: mgry_add/sub/mul/sqr/shift_left, DBLU, ZADDU,
                            ZDAU, ADD_Z2_1, TRPLU, scalar_mult, scalar_mult_p256,
                            from_affine / to_affine;
  * `ecsimd_b200.device` -- the same calls on device buffers (torch tensors are
                            used only to own device memory and streams).
"""
from . import capi  # noqa: F401
from .capi import Ecb200Error, device_count, init, init_devices, launch_count, shutdown  # noqa: F401
