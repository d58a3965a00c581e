"""Host-memory front-end with the reference's operator names.

Mirrors, batched over n lanes, the interface of aguinet/ecsimd's hot path
(include/ecsimd/mgry_ops.h, curve_group.h, jacobian_curve_point.h,
lib/scalar_mult_p256.cpp).  Inputs and outputs are numpy arrays in host memory;
every call goes through the C ABI with ECB200_MEM_HOST, i.e. it includes the
host->device and device->host copies.

Array shapes per layout (dtype uint32, C-contiguous):
  lane  : values (n, 8); affine points (n, 16) = x|y; Jacobian points (n, 24) = X|Y|Z
  pack4 : the reference's wide<> packs, as raw words: values (n/4, 32);
          affine (n/4, 64); Jacobian (n/4, 96)   [u64 word = limb*4 + lane]
  soa   : planar: values (2, n, 4); affine (4, n, 4); Jacobian (6, n, 4)
"""
import numpy as np

from . import capi
from .capi import LAYOUTS, MEM_HOST, NO_QUIRK

NC_VALUE, NC_AFFINE, NC_JAC = 1, 2, 3


def _shape(layout, n, nc):
    if layout == "lane":
        return (n, 8 * nc)
    if layout == "pack4":
        return (n // 4, 32 * nc)
    return (2 * nc, n, 4)


def lanes_of(a, layout, nc):
    """number of lanes held by array `a`"""
    a = np.asarray(a)
    words = a.size
    assert words % (8 * nc) == 0, "array size is not a whole number of %d-coordinate lanes" % nc
    return words // (8 * nc)


def _in(a):
    a = np.ascontiguousarray(a, dtype=np.uint32)
    return a


def _flags(layout, quirk):
    return LAYOUTS[layout] | MEM_HOST | (0 if quirk else NO_QUIRK)


def _run(name, outs_nc, ins, nc_in0, layout, quirk, extra=()):
    ins = [_in(x) for x in ins]
    n = lanes_of(ins[0], layout, nc_in0)
    outs = [np.zeros(_shape(layout, n, nc), np.uint32) for nc in outs_nc]
    args = [capi._p(o) for o in outs] + [capi._p(i) for i in ins] + list(extra) + [n, _flags(layout, quirk), None]
    capi.call(name, *args)
    return outs[0] if len(outs) == 1 else tuple(outs)


# ---- Montgomery field ops: include/ecsimd/mgry_ops.h --------------------------------------
def mgry_add(a, b, layout="lane"): return _run("ecb200_mgry_add", [1], [a, b], 1, layout, True)
def mgry_sub(a, b, layout="lane"): return _run("ecb200_mgry_sub", [1], [a, b], 1, layout, True)
def mgry_mul(a, b, layout="lane"): return _run("ecb200_mgry_mul", [1], [a, b], 1, layout, True)
def mgry_sqr(a, layout="lane", quirk=True): return _run("ecb200_mgry_sqr", [1], [a], 1, layout, quirk)
def mgry_shift_left(a, count=1, layout="lane"): return _run("ecb200_mgry_shift_left", [1], [a], 1, layout, True, extra=(count,))
def opposite(a, layout="lane"): return _run("ecb200_gfp_opposite", [1], [a], 1, layout, True)
def from_classical(a, layout="lane"): return _run("ecb200_from_classical", [1], [a], 1, layout, True)
def to_classical(a, layout="lane"): return _run("ecb200_to_classical", [1], [a], 1, layout, True)
def inverse(a, layout="lane", quirk=True): return _run("ecb200_gfp_inverse", [1], [a], 1, layout, quirk)
def mgry_mul_chain(a, b, iters, layout="lane"): return _run("ecb200_mgry_mul_chain", [1], [a, b], 1, layout, True, extra=(iters,))


# ---- co-Z point ops: include/ecsimd/curve_group.h -------------------------------------------
def DBLU(P, layout="lane", quirk=True):
    """-> (P rewritten, 2P)   curve_group.h:64-87"""
    return _run("ecb200_dblu", [3, 3], [P], 3, layout, quirk)


def TRPLU(P, layout="lane", quirk=True):
    """-> (P rewritten, 3P)   curve_group.h:183-186"""
    return _run("ecb200_trplu", [3, 3], [P], 3, layout, quirk)


def ZADDU(P, O, layout="lane", quirk=True):
    """-> (P rewritten, P+O)  curve_group.h:91-116"""
    return _run("ecb200_zaddu", [3, 3], [P, O], 3, layout, quirk)


def ZDAU(P, Q, layout="lane", quirk=True):
    """-> (Q rewritten, 2P+Q) curve_group.h:120-153"""
    return _run("ecb200_zdau", [3, 3], [P, Q], 3, layout, quirk)


def ADD_Z2_1(A, B, layout="lane", quirk=True):
    """-> A+B with Z(B) == R  curve_group.h:155-179"""
    return _run("ecb200_add_z2_1", [3], [A, B], 3, layout, quirk)


# ---- scalar multiplication -------------------------------------------------------------------
def scalar_mult(k, P, layout="lane", quirk=True):
    """curve_group<curve_nist_p256>::scalar_mult(x, P)  curve_group.h:189-218"""
    k = _in(k)
    n = lanes_of(k, layout, 1)
    P = _in(P)
    assert lanes_of(P, layout, 3) == n
    out = np.zeros(_shape(layout, n, 3), np.uint32)
    capi.call("ecb200_scalar_mult_p256", capi._p(out), capi._p(k), capi._p(P), n, _flags(layout, quirk), None)
    return out


scalar_mult_p256 = scalar_mult  # lib/scalar_mult_p256.cpp:12-14


def scalar_mult_base(k, layout="lane", quirk=True, table=True):
    """scalar_mult(x, WJG()): the generator for every lane (curve_group.h:39-41).
    table=False runs the plain ladder instead of starting from the fixed-base table (same results)."""
    k = _in(k)
    n = lanes_of(k, layout, 1)
    out = np.zeros(_shape(layout, n, 3), np.uint32)
    capi.call("ecb200_scalar_mult_p256_base", capi._p(out), capi._p(k), n, _flags(layout, quirk) | (0 if table else 0x200), None)
    return out


def scalar_mult_1s(k1, P, layout="lane", quirk=True):
    """curve_group::scalar_mult_1s(x, P): one scalar, all lanes  curve_group.h:221-251"""
    k1 = np.ascontiguousarray(k1, dtype=np.uint32).reshape(8)
    P = _in(P)
    n = lanes_of(P, layout, 3)
    out = np.zeros(_shape(layout, n, 3), np.uint32)
    capi.call("ecb200_scalar_mult_p256_1s", capi._p(out), capi._p(k1), capi._p(P), n, _flags(layout, quirk), None)
    return out


def from_affine(xy, layout="lane"):
    """wide_jacobian_curve_point::from_affine  jacobian_curve_point.h:25-31"""
    return _run("ecb200_from_affine", [3], [xy], 2, layout, True)


def to_affine(J, layout="lane", quirk=True):
    """wide_jacobian_curve_point::to_affine  jacobian_curve_point.h:33-42"""
    return _run("ecb200_to_affine", [2], [J], 3, layout, quirk)


def scalar_mult_affine(k, P=None, layout="lane", quirk=True, table=True):
    """curve_group::scalar_mult(x, P).to_affine() in one call (benchs/curve_group.cpp:28-46); P=None: the generator"""
    k = _in(k)
    n = lanes_of(k, layout, 1)
    out = np.zeros(_shape(layout, n, 2), np.uint32)
    P = None if P is None else _in(P)   # keep the converted array alive for the duration of the call
    capi.call("ecb200_scalar_mult_p256_affine", capi._p(out), capi._p(k), capi._p(P), n, _flags(layout, quirk) | (0 if table else 0x200), None)
    return out


def from_x(x, layout="lane", quirk=True):
    """wide_curve_point::from_x: y = sqrt(x^3 - 3x + b)  curve_point_ops.h:12-22; -> (y, ok[n] uint8)"""
    x = _in(x)
    n = lanes_of(x, layout, 1)
    y = np.zeros(_shape(layout, n, 1), np.uint32)
    ok = np.zeros(n, np.uint8)
    capi.call("ecb200_from_x", capi._p(y), capi._p(ok), capi._p(x), n, _flags(layout, quirk), None)
    return y, ok


class GenericField:
    """the field layer for a run-time modulus (reference: templates over P; its tests use secp256k1)"""

    def __init__(self, p_int):
        self.p = np.array([(p_int >> (32 * i)) & 0xFFFFFFFF for i in range(8)], np.uint32)

    def _u(self, name, a, extra=(), quirk=True, wout=8):
        a = _in(a)
        n = a.shape[0]
        out = np.zeros((n, wout), np.uint32)
        capi.call(name, capi._p(out), capi._p(a), *extra, capi._p(self.p), n, _flags("lane", quirk), None)
        return out

    def _b(self, name, a, b):
        a, b = _in(a), _in(b)
        out = np.zeros_like(a)
        capi.call(name, capi._p(out), capi._p(a), capi._p(b), capi._p(self.p), a.shape[0], _flags("lane", True), None)
        return out

    def mod_add(self, a, b): return self._b("ecb200_gen_mod_add", a, b)
    def mod_sub(self, a, b): return self._b("ecb200_gen_mod_sub", a, b)
    def mgry_mul(self, a, b): return self._b("ecb200_gen_mgry_mul", a, b)
    def mod_shift_left_one(self, a): return self._u("ecb200_gen_mod_shift_left_one", a)
    def mgry_sqr(self, a, quirk=True): return self._u("ecb200_gen_mgry_sqr", a, quirk=quirk)
    def from_classical(self, a): return self._u("ecb200_gen_from_classical", a)
    def to_classical(self, a): return self._u("ecb200_gen_to_classical", a)
    def opposite(self, a): return self._u("ecb200_gen_opposite", a)

    def mgry_pow(self, a, e_int, quirk=True):
        e = np.array([(e_int >> (32 * i)) & 0xFFFFFFFF for i in range(8)], np.uint32)
        return self._u("ecb200_gen_mgry_pow", a, extra=(capi._p(e),), quirk=quirk)


def mul512(a, b):
    """mul(a, b): exact 512-bit product, (n, 16) words  (mul.h:150-158)"""
    a, b = _in(a), _in(b)
    out = np.zeros((a.shape[0], 16), np.uint32)
    capi.call("ecb200_mul512", capi._p(out), capi._p(a), capi._p(b), a.shape[0], _flags("lane", True), None)
    return out


def square512(a):
    """square(a) of the reference, with its lost carry (mul.h:214-221)"""
    a = _in(a)
    out = np.zeros((a.shape[0], 16), np.uint32)
    capi.call("ecb200_square512", capi._p(out), capi._p(a), a.shape[0], _flags("lane", True), None)
    return out


def convert_layout(a, nc, src="lane", dst="soa"):
    """re-lay n lanes x nc coordinates on the device (pack4 <-> soa <-> lane)"""
    a = _in(a)
    n = lanes_of(a, src, nc)
    out = np.zeros(_shape(dst, n, nc), np.uint32)
    capi.call("ecb200_convert_layout", capi._p(out), LAYOUTS[dst], capi._p(a), LAYOUTS[src], nc, n, MEM_HOST, None)
    return out


def bn_from_bytes_BE(b, nc=1, layout="lane"):
    """serialization.h:12-23: (n, 32*nc) big-endian bytes -> values"""
    b = np.ascontiguousarray(b, dtype=np.uint8)
    n = b.size // (32 * nc)
    out = np.zeros(_shape(layout, n, nc), np.uint32)
    capi.call("ecb200_bn_from_bytes_be", capi._p(out), capi._p(b), nc, n, LAYOUTS[layout] | MEM_HOST, None)
    return out


def bn_to_bytes_BE(a, nc=1, layout="lane"):
    """serialization.h:25-48: values -> (n, 32*nc) big-endian bytes"""
    a = _in(a)
    n = lanes_of(a, layout, nc)
    out = np.zeros((n, 32 * nc), np.uint8)
    capi.call("ecb200_bn_to_bytes_be", capi._p(out), capi._p(a), nc, n, LAYOUTS[layout] | MEM_HOST, None)
    return out


def synth_values(seed, start, n, kind, layout="lane"):
    out = np.zeros(_shape(layout, n, 1), np.uint32)
    capi.call("ecb200_synth_values", capi._p(out), seed, start, kind, n, LAYOUTS[layout] | MEM_HOST, None)
    return out


# ---- layout transposition on the host (numpy), for callers and tests --------------------------
def lane_to_pack4(a, nc):
    """(n, 8*nc) lane-major -> (n/4, 32*nc) reference packs"""
    a = np.ascontiguousarray(a, np.uint32)
    n = a.shape[0]
    assert n % 4 == 0
    v = a.view(np.uint64).reshape(n // 4, 4, nc, 4)         # pack, lane, coord, limb
    return np.ascontiguousarray(v.transpose(0, 2, 3, 1)).view(np.uint32).reshape(n // 4, 32 * nc)


def pack4_to_lane(a, nc):
    a = np.ascontiguousarray(a, np.uint32)
    npk = a.shape[0]
    v = a.view(np.uint64).reshape(npk, nc, 4, 4)            # pack, coord, limb, lane
    return np.ascontiguousarray(v.transpose(0, 3, 1, 2)).view(np.uint32).reshape(npk * 4, 8 * nc)


def lane_to_soa(a, nc):
    a = np.ascontiguousarray(a, np.uint32)
    n = a.shape[0]
    v = a.reshape(n, nc, 2, 4)                               # lane, coord, half, word
    return np.ascontiguousarray(v.transpose(1, 2, 0, 3)).reshape(2 * nc, n, 4)


def soa_to_lane(a, nc):
    a = np.ascontiguousarray(a, np.uint32)
    n = a.shape[1]
    v = a.reshape(nc, 2, n, 4)
    return np.ascontiguousarray(v.transpose(2, 0, 1, 3)).reshape(n, 8 * nc)
