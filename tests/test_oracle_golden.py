"""CPU: the oracle against known answers that do not come from the oracle itself:
  * the reference's own KATs (tests/curve_group.cpp, tests/curve_point.cpp),
  * a pure-Python big-integer model of the field and of the affine group law,
  * the committed fixtures in tests/golden/ produced by the compiled reference
    (tests/golden/make_golden.py)."""
import json
import os

import numpy as np
import pytest

import _libs
from _libs import GX_INT, GY_INT, P_INT, R_INT, to_ints, to_words

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


# ---- independent affine model (python ints) ------------------------------------------------
def affine_add(p, q):
    if p is None: return q
    if q is None: return p
    (x1, y1), (x2, y2) = p, q
    if x1 == x2 and (y1 + y2) % P_INT == 0: return None
    if p == q: lam = (3 * x1 * x1 - 3) * pow(2 * y1, -1, P_INT) % P_INT
    else: lam = (y2 - y1) * pow(x2 - x1, -1, P_INT) % P_INT
    x3 = (lam * lam - x1 - x2) % P_INT
    return x3, (lam * (x1 - x3) - y1) % P_INT


def affine_mul(k, p):
    r = None
    while k:
        if k & 1: r = affine_add(r, p)
        p = affine_add(p, p)
        k >>= 1
    return r


def G_jac(orc, n=1):
    G = np.concatenate([to_words([GX_INT]), to_words([GY_INT])], axis=1)
    return orc.from_affine(np.repeat(G, n, axis=0))


def aff_ints(a):
    return [(to_ints(r[:8])[0], to_ints(r[8:])[0]) for r in a]


# ---- reference KATs --------------------------------------------------------------------------
def test_kat_dblu_zaddu_zdau(orc):
    """tests/curve_group.cpp:38-94: 2G, 3G (ZADDU and TRPLU), 5G (ZDAU), co-Z invariants"""
    GJ = G_jac(orc)
    P1, D = orc.dblu(GJ)
    assert np.array_equal(P1[:, 16:], D[:, 16:])                       # same Z (:45)
    assert aff_ints(orc.to_affine(P1))[0] == (GX_INT, GY_INT)          # P unchanged as a point (:46)
    assert aff_ints(orc.to_affine(D))[0] == (
        0x7CF27B188D034F7E8A52380304B51AC3C08969E277F21B35A60B48FC47669978,
        0x07775510DB8ED040293D9AC69F7430DBBA7DADE63CE982299E04B79D227873D1)  # :50-51
    P2, T = orc.zaddu(P1, D)
    g3 = (0x5ECBE4D1A6330A44C8F7EF951D4BF165E6C6B721EFADA985FB41661BC6E7FD6C,
          0x8734640C4998FF7E374B06CE1A64A2ECD82AB036384FB83D9A79B127A27D5032)
    assert aff_ints(orc.to_affine(T))[0] == g3                          # :66-67
    assert np.array_equal(P2[:, 16:], T[:, 16:])
    _, T2 = orc.trplu(GJ)
    assert np.array_equal(T2, T)
    Q, F = orc.zdau(D, P1)                                             # 2*(2G) + G = 5G (:78-94)
    g5 = (0x51590B7A515140D2D784C85608668FDFEF8C82FD1F5BE52421554A0DC3D033ED,
          0xE0C17DA8904A727D8AE1BF36BF8A79260D012F00D4D80888D1D0BB44FDA16DA4)
    assert aff_ints(orc.to_affine(F))[0] == g5
    assert np.array_equal(Q[:, 16:], F[:, 16:])
    assert aff_ints(orc.to_affine(Q))[0] == (GX_INT, GY_INT)


def test_kat_scalar_mult(orc):
    """tests/curve_group.cpp:117-173 (k = 5, 0bc1b1f2..., 0a891cec... even) + python model"""
    ks = [5, 0x0BC1B1F28709DECB543D9677D2CC9942348F6B984DEFF409430740942FF38827,
          0x0A891CEC7F6B6F8E0F2B3F6CC9F5E51D0B1A7C2BF6B3F3E7C4D5A6B7C8D9BD80, 3, 4, 6, 7, 2**255 + 12345, _libs.N_INT + 5]
    out = orc.scalar_mult(to_words(ks), G_jac(orc, len(ks)))
    aff = aff_ints(orc.to_affine(out))
    for k, a in zip(ks, aff):
        assert a == affine_mul(k, (GX_INT, GY_INT)), hex(k)
    # the Jacobian/Montgomery pattern for k = 0bc1... listed in SURVEY.md section 8c (dumped from the reference)
    assert to_ints(out[1, :8])[0] == 0x4C315298415AA6FEE7A24142CA3D3E5687E9DD69C99C308AD361C4341445835A
    assert to_ints(out[1, 8:16])[0] == 0xAA6CF5B34EA4BA14E76680E918BC8E19A38E60F112C49E92341052FD47611328
    assert to_ints(out[1, 16:])[0] == 0xD5488A3F8E4AB4C9DE98A83A0F210FED2A47CA4224EAF4F73105386F504ECA20


def test_kat_survey_rows(orc):
    """(k, P) -> (X,Y,Z) rows produced by the reference during the survey"""
    rows = [
        (0, 0xe52aba5fba2cacb27f441fc0a02b39eb52e3522e60bc57a48e1d676b75fd2448, 0x6f8063218184a6fb20bfa55669420656852fbc8a4c38ee94308180d4c107318f,
         0xdaeee89759406bf5b7b95770a1cc654ca44f31d64b5ea38297e7ab3a70b76068, 0xc869ed479c0f47cf556e72683c62426e5eb1423457c0fa4d6540233600696dbc, 0),
        (1, 0x034a184a716932d97a8f70439a65302ea399d6cbbb195850fdb16717cfb7d342, 0x242f3f30daf429997154f7606d09d3574429a5d5cf621f359213699c23846ffe,
         0xc4e1b567774b9c382232733ceb76cb4004d27e56c1d225bd936e6904e59efc06, 0x3cf4f02a94c44fe7700c5b38e72c8e57b17871f271f5ad495756ed0f471b0c5a,
         0x2b342b07f9e0b82f75e35c279f0a72b099dbf340819f262c2f5e4e63fcd74fd9),
        (0xed8702f78af0241393530d6c08116507dcc752fbc4ccb905ce9690136e1287f6, 0x348598bfc9582058d1187326b494241d91ccfe5cb1d1edbb4f932b72feb1916c,
         0x63b18bdf13093fc3f3488bb1afbca2269987bfc4d366354db93b4b587ee61fc9, 0x6686a504f8be66038c4be35ac69a1f0dfb8e6480fce5192ce99100d860bae836,
         0x46fec89308ca3afe3bc1f13dd680718a20e0e8bcfa1866d02a7904c4c7e7ba66, 0x7755eb17610b89f4294d609c4f18fc162dd808352d3e3ffd056d24af10b0bab8)]
    k = to_words([r[0] for r in rows])
    P = np.concatenate([to_words([r[1] for r in rows]), to_words([r[2] for r in rows]), to_words([R_INT % P_INT] * len(rows))], axis=1)
    out = orc.scalar_mult(k, P)
    for i, r in enumerate(rows):
        assert (to_ints(out[i, :8])[0], to_ints(out[i, 8:16])[0], to_ints(out[i, 16:])[0]) == r[3:], i


def test_kat_from_x(orc):
    """tests/curve_point.cpp:17-26: decompression of Gx gives Gy or p-Gy"""
    import ctypes as C
    x = to_words([GX_INT])
    y = np.zeros((1, 8), np.uint32); ok = np.zeros(1, np.uint8)
    f = orc.lib.orc_from_x; f.restype = None
    f(y.ctypes.data_as(C.c_void_p), ok.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p), C.c_size_t(1), C.c_int(1))
    assert ok[0] == 1 and to_ints(y)[0] in (GY_INT, P_INT - GY_INT)


# ---- python big-int model of the field layer ---------------------------------------------------
def test_field_vs_python_model(orc):
    n = 2000
    a, b = _libs.field_elems(101, n), _libs.field_elems(102, n)
    ai, bi = to_ints(a), to_ints(b)
    Rinv = pow(R_INT, -1, P_INT)
    assert to_ints(orc.mgry_add(a, b)) == [(x + y) % P_INT for x, y in zip(ai, bi)]
    assert to_ints(orc.mgry_sub(a, b)) == [(x - y) % P_INT for x, y in zip(ai, bi)]
    assert to_ints(orc.mgry_mul(a, b)) == [x * y * Rinv % P_INT for x, y in zip(ai, bi)]
    assert to_ints(orc.mgry_shl1(a)) == [2 * x % P_INT for x in ai]
    assert to_ints(orc.opposite(a)) == [(-x) % P_INT for x in ai]
    assert to_ints(orc.from_classical(a)) == [x * R_INT % P_INT for x in ai]
    assert to_ints(orc.to_classical(a)) == [x * Rinv % P_INT for x in ai]
    # squares: equal to the true square except on quirk lanes (none expected in 2000 random draws)
    assert to_ints(orc.mgry_sqr(a)) == [x * x * Rinv % P_INT for x in ai]


def test_quirk_vectors(orc):
    """SURVEY.md 8a-Q pinned vectors: reference square() != a^2"""
    a = to_words([0x196E98832350A697302E3812CF37CFDB65BD91769E220A413C2BC6519E220A41])
    sq = orc.square512(a)[0]
    got = sum(int(w) << (32 * i) for i, w in enumerate(sq))
    want = int("0286c991087416f5a0939d93f099dfdfd51866b166ab118037fff802a7cf7a58"
               "31f47a9b945f36517986f454ab1179b139b6bdde6687b7e204eb1250f5ad2481", 16)
    assert got == want
    true_sq = to_ints(a)[0] ** 2
    assert got != true_sq and true_sq - got == 1 << 256        # exactly one lost carry
    b = to_words([0xA09D838E868B90F2B89CC416F270D3B8F0374E0A8728A79978B896A45AF4F8A8])
    s, m = to_ints(orc.mgry_sqr(b))[0], to_ints(orc.mgry_mul(b, b))[0]
    assert s != m and (s & 0xFFFFFFFF) == 0xA47077CB and (m & 0xFFFFFFFF) == 0xA47077CC


def test_quirk_filter_is_sound(orc):
    """every lane where square() loses a carry has a cross product with high word 0x7fffffff
    (the necessary condition the CUDA fast path tests, csrc/fp256.cuh)"""
    x = _libs.quirk_stress(20000, seed=11)
    q = (orc.mgry_sqr(x) != orc.mgry_mul(x, x)).any(axis=1)
    xs = x.astype(np.uint64)
    hit = np.zeros(len(x), bool)
    for i in range(8):
        for j in range(i + 1, 8):
            hit |= ((xs[:, i] * xs[:, j]) >> np.uint64(32)) == np.uint64(0x7FFFFFFF)
    assert q.sum() > 1000 and not (q & ~hit).any()


def _fp32_filter_bits(ai, aj):
    """numpy model of the CUDA first-level filter (csrc/fp256.cuh, fp_sqr_quirk_filter):
    f(a) = as_float(0x3f000000 + (a >> 8)); bits(fma(-f_i, f_j, 2.0f)).  f_i*f_j (48 significant bits)
    and 2 - f_i*f_j are exact in float64, so one float64 -> float32 conversion is the fma's rounding."""
    fi = ((ai >> np.uint64(8)).astype(np.uint32) + np.uint32(0x3F000000)).view(np.float32).astype(np.float64)
    fj = ((aj >> np.uint64(8)).astype(np.uint32) + np.uint32(0x3F000000)).view(np.float32).astype(np.float64)
    return (2.0 - fi * fj).astype(np.float32).view(np.uint32)


def test_fp32_first_level_filter_is_sound():
    """no false negative: every 32-bit pair whose product has high word 0x7fffffff passes the fp32 test
    (bits <= 0x35000000), on pairs packed against both ends of every admissible product range"""
    rnd = np.random.RandomState(12)
    ai = (rnd.randint(0, 1 << 31, size=2_000_000).astype(np.uint64) | np.uint64(1 << 31))
    ai[:64] = np.uint64(1 << 31); ai[64:128] = np.uint64(0xFFFFFFFF); ai[128:192] = np.uint64(0xB504F333)   # 2^31, 2^32-1, ~2^31.5
    lo, hi = (np.uint64((1 << 63) - (1 << 32)) + ai - np.uint64(1)) // ai, np.uint64((1 << 63) - 1) // ai    # a_j range of a hit
    ok = (lo <= hi) & (hi < np.uint64(1 << 32))
    ai, lo, hi = ai[ok], lo[ok], hi[ok]
    assert len(ai) > 1_000_000
    for aj in (lo, hi):
        assert (((ai * aj) >> np.uint64(32)) == np.uint64(0x7FFFFFFF)).all()
        assert (_fp32_filter_bits(ai, aj) <= np.uint32(0x35000000)).all()
        assert (_fp32_filter_bits(aj, ai) <= np.uint32(0x35000000)).all()
    # and it is selective: random pairs almost never pass (the 16-bit integer form passed 3e-5 of them)
    x = rnd.randint(0, 1 << 32, size=4_000_000, dtype=np.uint64)
    y = rnd.randint(0, 1 << 32, size=4_000_000, dtype=np.uint64)
    assert (_fp32_filter_bits(x, y) <= np.uint32(0x35000000)).mean() < 2e-6


# ---- committed fixtures generated from the compiled reference -----------------------------------
def _load(name):
    path = os.path.join(GOLDEN, name)
    if not os.path.exists(path):
        pytest.skip("fixture %s missing" % name)
    return np.load(path)


def test_golden_field(orc):
    g = _load("field_ops.npz")
    a, b = g["a"], g["b"]
    for op in ("mgry_add", "mgry_sub", "mgry_mul"):
        assert np.array_equal(getattr(orc, op)(a, b), g[op]), op
    for op in ("mgry_sqr", "mgry_shl1", "opposite", "from_classical", "to_classical"):
        assert np.array_equal(getattr(orc, op)(a), g[op]), op
    assert np.array_equal(orc.mul512(a, b), g["mul512"])
    assert np.array_equal(orc.square512(a), g["square512"])


def test_golden_points(orc):
    g = _load("point_ops.npz")
    P, k = g["P"], g["k"]
    p1, d = orc.dblu(P)
    assert np.array_equal(p1, g["dblu_P"]) and np.array_equal(d, g["dblu_2P"])
    p2, t = orc.zaddu(p1, d)
    assert np.array_equal(p2, g["zaddu_P"]) and np.array_equal(t, g["zaddu_R"])
    q, r = orc.zdau(t, p2)
    assert np.array_equal(q, g["zdau_Q"]) and np.array_equal(r, g["zdau_R"])
    assert np.array_equal(orc.add_z2_1(r, P), g["add_z2_1"])
    assert np.array_equal(orc.scalar_mult(k, P), g["scalar_mult"])
    assert np.array_equal(orc.to_affine(g["scalar_mult"]), g["to_affine"])


def test_golden_quirk_scalar_mult(orc):
    path = os.path.join(GOLDEN, "quirk_scalar_mult.json")
    if not os.path.exists(path):
        pytest.skip("no quirk scalar-mult fixture")
    fx = json.load(open(path))
    k = to_words([int(v, 16) for v in fx["k"]])
    P = np.concatenate([to_words([int(v, 16) for v in fx["Px"]]), to_words([int(v, 16) for v in fx["Py"]]),
                        to_words([R_INT % P_INT] * len(fx["k"]))], axis=1)
    want = np.concatenate([to_words([int(v, 16) for v in fx[c]]) for c in ("X", "Y", "Z")], axis=1)
    assert np.array_equal(orc.scalar_mult(k, P), want)
