"""GPU parity: Montgomery field kernels (C ABI, host buffers) vs the CPU oracle.

Bit-exact on canonical, edge, squaring-quirk AND arbitrary non-canonical inputs
(the reference's functions are deterministic on any 256-bit pattern; so are ours).
Reference tests mirrored: tests/mgry.cpp (Mgry.FromTo, Mgry.Ops, Mgry.Gfp),
tests/ops.cpp (Ops256.Mod), re-stated for the P-256 prime through the oracle.
"""
import numpy as np
import pytest

import _libs
from _libs import EDGE_FIELD, QUIRK_FIELD, field_elems, raw256, to_words

pytestmark = pytest.mark.gpu


def _inputs(n, seed):
    a = field_elems(seed, n)
    b = field_elems(seed + 1, n)
    edge = to_words(EDGE_FIELD + QUIRK_FIELD)
    # all pairs of edge values
    ea = np.repeat(edge, len(edge), axis=0)
    eb = np.tile(edge, (len(edge), 1))
    return np.concatenate([ea, a]), np.concatenate([eb, b])


@pytest.mark.parametrize("op", ["mgry_add", "mgry_sub", "mgry_mul"])
def test_binary_ops_canonical(eng, orc, op):
    a, b = _inputs(1 << 14, 0xEC51D001)
    assert np.array_equal(getattr(eng, op)(a, b), getattr(orc, op)(a, b))


@pytest.mark.parametrize("op", ["mgry_add", "mgry_sub", "mgry_mul"])
def test_binary_ops_any_bit_pattern(eng, orc, op):
    a, b = raw256(11, 1 << 13), raw256(12, 1 << 13)
    a[:512, 7] = 0xFFFFFFFF; b[:256, 7] = 0xFFFFFFFF; a[:128, 6] = 0xFFFFFFFF; b[100:300, 6] = 0xFFFFFFFF
    a[300:400] = 0xFFFFFFFF; b[350:450] = 0xFFFFFFFF
    assert np.array_equal(getattr(eng, op)(a, b), getattr(orc, op)(a, b))


def test_unary_ops(eng, orc):
    a, _ = _inputs(1 << 13, 0xEC51D002)
    r = raw256(13, 1 << 12); r[:256, 7] = 0xFFFFFFFF; r[:64] = 0xFFFFFFFF
    for x in (a, r):
        assert np.array_equal(eng.mgry_sqr(x), orc.mgry_sqr(x))
        assert np.array_equal(eng.mgry_shift_left(x, 1), orc.mgry_shl1(x))
        assert np.array_equal(eng.opposite(x), orc.opposite(x))
        assert np.array_equal(eng.from_classical(x), orc.from_classical(x))
        assert np.array_equal(eng.to_classical(x), orc.to_classical(x))
    s3 = orc.mgry_shl1(orc.mgry_shl1(orc.mgry_shl1(a)))
    assert np.array_equal(eng.mgry_shift_left(a, 3), s3)


def test_from_to_classical_roundtrip(eng):
    a = field_elems(5, 1 << 14)
    assert np.array_equal(eng.to_classical(eng.from_classical(a)), a)


def test_sqr_quirk_vectors(eng, orc):
    """the reference's square() loses a carry on these; we must reproduce it"""
    q = to_words(QUIRK_FIELD)
    got = eng.mgry_sqr(q)
    assert np.array_equal(got, orc.mgry_sqr(q))
    assert not np.array_equal(got, orc.mgry_mul(q, q))          # it really is the defect
    assert np.array_equal(eng.mgry_sqr(q, quirk=False), orc.mgry_mul(q, q))  # and NO_QUIRK is the true square


def test_sqr_quirk_stress(eng, orc):
    """inputs built so that some cross product a_i*a_j sits just below 2^63: dense in filter hits"""
    x = _libs.quirk_stress(20000, seed=7)
    want = orc.mgry_sqr(x)
    got = eng.mgry_sqr(x)
    assert np.array_equal(got, want)
    true_sq = orc.mgry_mul(x, x)
    assert (want != true_sq).any(axis=1).sum() > 100   # the set does exercise the defect


def test_inverse(eng, orc):
    a = field_elems(9, 512)
    a[0] = to_words([1])[0]; a[1] = to_words([_libs.P_INT - 1])[0]
    inv = eng.inverse(a)
    assert np.array_equal(inv, orc.inverse(a))
    R = to_words([_libs.R_INT % _libs.P_INT])
    assert np.array_equal(eng.mgry_mul(inv, a), np.repeat(R, 512, axis=0))


@pytest.mark.parametrize("layout", ["pack4", "soa"])
def test_layouts(eng, orc, layout):
    n = 4096
    a, b = field_elems(21, n), field_elems(22, n)
    conv = {"pack4": (eng.lane_to_pack4, eng.pack4_to_lane), "soa": (eng.lane_to_soa, eng.soa_to_lane)}[layout]
    for op in ("mgry_add", "mgry_sub", "mgry_mul"):
        got = conv[1](getattr(eng, op)(conv[0](a, 1), conv[0](b, 1), layout=layout), 1)
        assert np.array_equal(got, getattr(orc, op)(a, b))
    got = conv[1](eng.mgry_sqr(conv[0](a, 1), layout=layout), 1)
    assert np.array_equal(got, orc.mgry_sqr(a))


def test_ragged_and_empty(eng, orc):
    for n in (0, 1, 3, 31, 33, 255, 257):
        a, b = field_elems(31, n), field_elems(32, n)
        got = eng.mgry_mul(a, b)
        assert got.shape == (n, 8)
        if n:
            assert np.array_equal(got, orc.mgry_mul(a, b))


def test_mul_chain_matches_repeated_mul(eng, orc):
    a, b = field_elems(41, 1024), field_elems(42, 1024)
    want = a
    for _ in range(5):
        want = orc.mgry_mul(want, b)
    assert np.array_equal(eng.mgry_mul_chain(a, b, 5), want)


def test_full_size_properties(eng):
    """config 1 at full size (2^20): linearity/commutativity properties, oracle on a sample"""
    n = 1 << 20
    a, b, c = field_elems(0xEC51D001, n), field_elems(0xEC51D002, n), field_elems(0xEC51D005, n)
    ab = eng.mgry_mul(a, b)
    assert np.array_equal(ab, eng.mgry_mul(b, a))
    # (a+b)*c == a*c + b*c
    lhs = eng.mgry_mul(eng.mgry_add(a, b), c)
    rhs = eng.mgry_add(eng.mgry_mul(a, c), eng.mgry_mul(b, c))
    assert np.array_equal(lhs, rhs)
    # a - b + b == a
    assert np.array_equal(eng.mgry_add(eng.mgry_sub(a, b), b), a)
    o = _libs.oracle(8)
    idx = np.arange(0, n, 257)
    assert np.array_equal(ab[idx], o.mgry_mul(a[idx], b[idx]))


def test_layout_and_byte_adapters(eng):
    """pack4 <-> soa <-> lane on the device equals the numpy transposition; BE byte strings round-trip
    and equal int.to_bytes (serialization.h:12-48 of the reference)"""
    n = 1024
    a = raw256(71, 3 * n).reshape(n, 24)
    for src, dst in (("lane", "soa"), ("lane", "pack4"), ("pack4", "soa"), ("soa", "pack4"), ("soa", "lane"), ("pack4", "lane")):
        conv = {"lane": lambda x: x, "soa": lambda x: eng.lane_to_soa(x, 3), "pack4": lambda x: eng.lane_to_pack4(x, 3)}
        assert np.array_equal(eng.convert_layout(conv[src](a), 3, src, dst), conv[dst](a)), (src, dst)
    v = raw256(72, 2 * n).reshape(n, 16)
    b = eng.bn_to_bytes_BE(v, nc=2)
    ints = _libs.to_ints(v.reshape(-1, 8))
    want = np.frombuffer(b"".join(x.to_bytes(32, "big") for x in ints), np.uint8).reshape(n, 64)
    assert np.array_equal(b, want)
    assert np.array_equal(eng.bn_from_bytes_BE(b, nc=2), v)
    assert np.array_equal(eng.soa_to_lane(eng.bn_from_bytes_BE(b, nc=2, layout="soa"), 2), v)
