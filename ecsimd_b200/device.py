"""Device-memory front-end: the same C-ABI calls on buffers that already live in HBM.

torch is used only to own device memory and CUDA streams (plumbing); buffers are
uint32 tensors in the planar ECB200_LAYOUT_SOA layout unless stated otherwise:
values (2, n, 4), affine points (4, n, 4), Jacobian points (6, n, 4).
Calls are asynchronous on torch's current stream.
"""
import torch

from . import capi
from .capi import LAYOUTS, MEM_DEVICE, NO_QUIRK


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _flags(layout, quirk=True):
    return LAYOUTS[layout] | MEM_DEVICE | (0 if quirk else NO_QUIRK)


def empty(n, nc, layout="soa", device=None):
    shape = {"soa": (2 * nc, n, 4), "lane": (n, 8 * nc), "pack4": (max(n // 4, 1), 32 * nc)}[layout]
    return torch.empty(shape, dtype=torch.int32, device=device or torch.device("cuda", torch.cuda.current_device()))


def synth_values(out, seed, start, n, kind, layout="soa"):
    capi.call("ecb200_synth_values", out.data_ptr(), seed, start, kind, n, _flags(layout), _stream())
    return out


def mgry_mul(out, a, b, n, layout="soa"):
    capi.call("ecb200_mgry_mul", out.data_ptr(), a.data_ptr(), b.data_ptr(), n, _flags(layout), _stream())
    return out


def mgry_sqr(out, a, n, layout="soa", quirk=True):
    capi.call("ecb200_mgry_sqr", out.data_ptr(), a.data_ptr(), n, _flags(layout, quirk), _stream())
    return out


def mgry_add(out, a, b, n, layout="soa"):
    capi.call("ecb200_mgry_add", out.data_ptr(), a.data_ptr(), b.data_ptr(), n, _flags(layout), _stream())
    return out


def mgry_sub(out, a, b, n, layout="soa"):
    capi.call("ecb200_mgry_sub", out.data_ptr(), a.data_ptr(), b.data_ptr(), n, _flags(layout), _stream())
    return out


def mgry_mul_chain(out, a, b, iters, n, layout="soa"):
    capi.call("ecb200_mgry_mul_chain", out.data_ptr(), a.data_ptr(), b.data_ptr(), iters, n, _flags(layout), _stream())
    return out


def scalar_mult(out, k, P, n, layout="soa", quirk=True):
    capi.call("ecb200_scalar_mult_p256", out.data_ptr(), k.data_ptr(), P.data_ptr(), n, _flags(layout, quirk), _stream())
    return out


def scalar_mult_base(out, k, n, layout="soa", quirk=True, table=True):
    capi.call("ecb200_scalar_mult_p256_base", out.data_ptr(), k.data_ptr(), n, _flags(layout, quirk) | (0 if table else 0x200), _stream())
    return out


def trplu(outP, out3, P, n, layout="soa", quirk=True):
    capi.call("ecb200_trplu", outP.data_ptr(), out3.data_ptr(), P.data_ptr(), n, _flags(layout, quirk), _stream())


def dblu(outP, out2, P, n, layout="soa", quirk=True):
    capi.call("ecb200_dblu", outP.data_ptr(), out2.data_ptr(), P.data_ptr(), n, _flags(layout, quirk), _stream())


def zaddu(outP, outR, P, O, n, layout="soa", quirk=True):
    capi.call("ecb200_zaddu", outP.data_ptr(), outR.data_ptr(), P.data_ptr(), O.data_ptr(), n, _flags(layout, quirk), _stream())


def add_z2_1(outR, A, B, n, layout="soa", quirk=True):
    capi.call("ecb200_add_z2_1", outR.data_ptr(), A.data_ptr(), B.data_ptr(), n, _flags(layout, quirk), _stream())


def zdau(outQ, outR, P, Q, n, layout="soa", quirk=True):
    capi.call("ecb200_zdau", outQ.data_ptr(), outR.data_ptr(), P.data_ptr(), Q.data_ptr(), n, _flags(layout, quirk), _stream())


def to_affine(xy, J, n, layout="soa", quirk=True):
    capi.call("ecb200_to_affine", xy.data_ptr(), J.data_ptr(), n, _flags(layout, quirk), _stream())
    return xy


def scalar_mult_affine(xy, k, P, n, layout="soa", quirk=True):
    capi.call("ecb200_scalar_mult_p256_affine", xy.data_ptr(), k.data_ptr(), None if P is None else P.data_ptr(), n, _flags(layout, quirk), _stream())
    return xy


def inverse(out, a, n, layout="soa", quirk=True):
    capi.call("ecb200_gfp_inverse", out.data_ptr(), a.data_ptr(), n, _flags(layout, quirk), _stream())
    return out


def from_x(y, ok, x, n, layout="soa", quirk=True):
    """y: (2, n, 4) classical y; ok: (n,) uint8 device tensor; x: classical x"""
    capi.call("ecb200_from_x", y.data_ptr(), ok.data_ptr(), x.data_ptr(), n, _flags(layout, quirk), _stream())
    return y, ok


def from_affine(J, xy, n, layout="soa"):
    capi.call("ecb200_from_affine", J.data_ptr(), xy.data_ptr(), n, _flags(layout), _stream())
    return J


def checksum(buf):
    import numpy as np
    out = np.zeros(8, np.uint32)
    capi.call("ecb200_checksum", capi._p(out), buf.data_ptr(), buf.numel(), _stream())
    return out


def microbench(which, blocks, threads, iters):
    import ctypes as C
    ops = C.c_double()
    ms = C.c_float()
    capi.call("ecb200_microbench", which, blocks, threads, iters, C.byref(ops), C.byref(ms), _stream())
    total = ops.value * iters * blocks * threads
    return total / (ms.value * 1e-3), ms.value
