"""CPU: pins the C oracle (oracle/p256_oracle.c) bit-for-bit to the reference's own code
compiled from /root/reference (oracle/_ref/libecsimd_ref.so).  Skipped where that
library is absent (e.g. on the GPU box if the snapshot did not carry it); the
committed fixtures of test_oracle_golden.py cover that case."""
import ctypes as C

import numpy as np
import pytest

import _libs
from _libs import EDGE_FIELD, EDGE_SCALARS, GX_INT, GY_INT, QUIRK_FIELD, field_elems, raw256, to_words


def _mix(n, seed):
    edge = to_words(EDGE_FIELD + QUIRK_FIELD)
    a = np.concatenate([np.repeat(edge, len(edge), axis=0), field_elems(seed, n), _libs.quirk_stress(256, seed)])
    b = np.concatenate([np.tile(edge, (len(edge), 1)), field_elems(seed + 1, n), field_elems(seed + 2, 256)])
    return a, b


def test_constants(orc, ref):
    """P, R, R^2, (p-1)R, Am, Bm, G (Montgomery): mgry_csts.h:20-24, curve_group.h:31-41"""
    assert np.array_equal(orc.constants(), ref.constants())


def test_pack_layout(ref):
    """word index inside a wide<bignum_256> pack is limb*4 + lane; sizes 128/384/256 bytes"""
    lanes = np.arange(16, dtype=np.uint64).reshape(4, 4) + np.uint64(100)   # [lane][limb]
    out = np.zeros(16, np.uint64)
    f = ref.lib.ref_pack_layout_probe; f.restype = None
    f(out.ctypes.data_as(C.c_void_p), lanes.ctypes.data_as(C.c_void_p))
    assert np.array_equal(out.reshape(4, 4), lanes.T)                       # [limb][lane]
    for name, size in (("ref_sizeof_wbn", 128), ("ref_sizeof_wjcp", 384), ("ref_sizeof_wcp", 256)):
        g = getattr(ref.lib, name); g.restype = C.c_size_t
        assert g() == size
    # and the host-side transposition helper agrees with it
    from ecsimd_b200 import host
    a = field_elems(3, 4)
    pk = host.lane_to_pack4(a, 1).view(np.uint64).reshape(4, 4)
    assert np.array_equal(pk, a.view(np.uint64).reshape(4, 4).T)


@pytest.mark.parametrize("op", ["mgry_add", "mgry_sub", "mgry_mul"])
def test_binary(orc, ref, op):
    a, b = _mix(4096, 1)
    assert np.array_equal(getattr(orc, op)(a, b), getattr(ref, op)(a, b))
    a, b = raw256(3, 4096), raw256(4, 4096)           # any bit pattern, incl. >= p
    a[:256, 7] = 0xFFFFFFFF; b[:128, 7] = 0xFFFFFFFF; a[:64] = 0xFFFFFFFF
    assert np.array_equal(getattr(orc, op)(a, b), getattr(ref, op)(a, b))


@pytest.mark.parametrize("op", ["mgry_shl1", "mgry_sqr", "opposite", "from_classical", "to_classical"])
def test_unary(orc, ref, op):
    a, _ = _mix(4096, 5)
    assert np.array_equal(getattr(orc, op)(a), getattr(ref, op)(a))
    r = raw256(6, 2048); r[:256, 7] = 0xFFFFFFFF; r[:32] = 0xFFFFFFFF
    assert np.array_equal(getattr(orc, op)(r), getattr(ref, op)(r))


def test_integer_layer(orc, ref):
    """mul (mul.h:150-158), square with its defect (mul.h:160-221), mgry_reduce (mgry_mul.h:84-121)"""
    a, b = _mix(2048, 9)
    assert np.array_equal(orc.mul512(a, b), ref.mul512(a, b))
    assert np.array_equal(orc.square512(a), ref.square512(a))
    t = np.concatenate([raw256(10, 4096), raw256(11, 4096)], axis=1)   # arbitrary 512-bit inputs
    t[:64] = 0xFFFFFFFF
    assert np.array_equal(orc.mgry_reduce(t), ref.mgry_reduce(t))


def test_quirk_stress(orc, ref):
    x = _libs.quirk_stress(20000, seed=3)
    assert np.array_equal(orc.mgry_sqr(x), ref.mgry_sqr(x))
    assert np.array_equal(orc.square512(x), ref.square512(x))


def test_inverse(orc, ref):
    a = field_elems(12, 64)
    assert np.array_equal(orc.inverse(a), ref.inverse(a))


def _pts(ref, n, seed):
    G = np.concatenate([to_words([GX_INT]), to_words([GY_INT])], axis=1)
    GJ = ref.from_affine(np.repeat(G, n, axis=0))
    return ref.from_affine(ref.to_affine(ref.scalar_mult(raw256(seed, n), GJ)))


def test_point_ops(orc, ref):
    P = _pts(ref, 128, 20)
    a1, a2 = ref.dblu(P); b1, b2 = orc.dblu(P)
    assert np.array_equal(a1, b1) and np.array_equal(a2, b2)
    c1, c2 = ref.zaddu(a1, a2); d1, d2 = orc.zaddu(a1, a2)
    assert np.array_equal(c1, d1) and np.array_equal(c2, d2)
    e1, e2 = ref.zdau(c2, c1); f1, f2 = orc.zdau(c2, c1)
    assert np.array_equal(e1, f1) and np.array_equal(e2, f2)
    g1, g2 = ref.trplu(P); h1, h2 = orc.trplu(P)
    assert np.array_equal(g1, h1) and np.array_equal(g2, h2)
    assert np.array_equal(ref.add_z2_1(e2, P), orc.add_z2_1(e2, P))
    assert np.array_equal(ref.to_affine(e2), orc.to_affine(e2))
    assert np.array_equal(ref.from_affine(ref.to_affine(e2)), orc.from_affine(orc.to_affine(e2)))
    # garbage in, same garbage out
    X = raw256(21, 3 * 64).reshape(64, 24); Y = raw256(22, 3 * 64).reshape(64, 24)
    e1, e2 = ref.zdau(X, Y); f1, f2 = orc.zdau(X, Y)
    assert np.array_equal(e1, f1) and np.array_equal(e2, f2)


def test_scalar_mult(orc, ref):
    n = 256
    P = _pts(ref, n, 30)
    k = raw256(31, n)
    for i, v in enumerate(EDGE_SCALARS):
        k[i] = to_words([v])[0]
    want = ref.scalar_mult(k, P)
    assert np.array_equal(orc.scalar_mult(k, P), want)
    # ragged tail (n not a multiple of the 4-lane pack)
    assert np.array_equal(orc.scalar_mult(k[:7], P[:7]), ref.scalar_mult(k[:7], P[:7]))
    # scalar_mult_1s gives the same lanes as scalar_mult with a broadcast scalar (tests/curve_group.cpp:130-139)
    f = ref.lib.ref_scalar_mult_1s; f.restype = None
    out = np.zeros((n, 24), np.uint32)
    k1 = np.ascontiguousarray(k[20])
    f(out.ctypes.data_as(C.c_void_p), k1.ctypes.data_as(C.c_void_p), P.ctypes.data_as(C.c_void_p), C.c_size_t(n), C.c_int(4))
    assert np.array_equal(out, orc.scalar_mult(np.repeat(k1[None], n, axis=0), P))


def test_from_x(orc, ref):
    n = 64
    P = _pts(ref, n, 40)
    x = ref.to_affine(P)[:, :8].copy()
    y_r = np.zeros((n, 8), np.uint32); ok_r = np.zeros(n // 4, np.uint8)
    f = ref.lib.ref_from_x; f.restype = None
    f(y_r.ctypes.data_as(C.c_void_p), ok_r.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p), C.c_size_t(n))
    y_o = np.zeros((n, 8), np.uint32); ok_o = np.zeros(n, np.uint8)
    g = orc.lib.orc_from_x; g.restype = None
    g(y_o.ctypes.data_as(C.c_void_p), ok_o.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p), C.c_size_t(n), C.c_int(2))
    assert ok_r.all() and ok_o.all() and np.array_equal(y_r, y_o)
