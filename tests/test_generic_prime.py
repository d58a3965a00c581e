"""The field layer on a run-time modulus.  CPU: the oracle's generic functions against the
reference compiled for the secp256k1 prime and against the reference's own KATs (tests/mgry.cpp
Mgry.FromTo/Ops/Gfp, tests/ops.cpp Ops256.Binops/Mod).  GPU: the engine against the oracle."""
import ctypes as C

import numpy as np
import pytest

import _libs
from _libs import raw256, to_ints, to_words

K1 = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEFFFFFC2F
OPS = {"mod_add": 0, "mod_sub": 1, "mod_shift_left_one": 2, "mgry_mul": 3, "mgry_sqr": 4, "from_classical": 5,
       "to_classical": 6, "mgry_pow": 7, "opposite": 8}


def _p(a):
    return None if a is None else np.ascontiguousarray(a, np.uint32).ctypes.data_as(C.c_void_p)


def orc_op(orc, name, a, b=None, e=None, p=K1):
    a = np.ascontiguousarray(a, np.uint32)
    o = np.zeros_like(a)
    f = orc.lib.orc_gen_op; f.restype = None
    f(C.c_int(OPS[name]), _p(o), _p(a), _p(b), _p(to_words([e])[0]) if e is not None else None, _p(to_words([p])[0]), C.c_size_t(a.shape[0]))
    return o


def ref_op(ref, name, a, b=None, e=None):
    a = np.ascontiguousarray(a, np.uint32)
    o = np.zeros_like(a)
    f = ref.lib.ref_k1_op; f.restype = None
    f(C.c_int(OPS[name]), _p(o), _p(a), _p(b), _p(to_words([e])[0]) if e is not None else None, C.c_size_t(a.shape[0]))
    return o


def _vals(n, seed):
    a = raw256(seed, n)
    a[: n // 8, 7] = 0xFFFFFFFF; a[: n // 16] = 0xFFFFFFFF          # values >= p and all-ones
    a[n // 2:] = to_words([v % K1 for v in to_ints(a[n // 2:])])      # canonical half
    return a


def test_oracle_generic_vs_reference_secp256k1(orc, ref):
    a, b = _vals(1024, 1), _vals(1024, 2)
    for name in ("mod_add", "mod_sub", "mgry_mul"):
        assert np.array_equal(orc_op(orc, name, a, b), ref_op(ref, name, a, b)), name
    for name in ("mod_shift_left_one", "mgry_sqr", "from_classical", "to_classical", "opposite"):
        assert np.array_equal(orc_op(orc, name, a), ref_op(ref, name, a)), name
    q = _libs.quirk_stress(2000, 3)
    assert np.array_equal(orc_op(orc, "mgry_sqr", q), ref_op(ref, "mgry_sqr", q))
    for e in (K1 - 2, (K1 + 1) // 4, 2, 0, 1, 0x00000000000F0000000000000000000000000000000000000000000000000001):
        assert np.array_equal(orc_op(orc, "mgry_pow", a[:64], e=e), ref_op(ref, "mgry_pow", a[:64], e=e)), hex(e)


def _kats(op):
    """the reference's KATs, op = callable(name, a, b=None, e=None) on (n,8) word arrays"""
    h = lambda s: to_words([int(s, 16)])
    # Mgry.FromTo  tests/mgry.cpp:32-50
    vals = to_words([int(x, 16) for x in (
        "eeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeee", "0168db3a8eca3fd7d4d08943182e189aef318068ba8853d77cb49c17bae00c0e",
        "2714dac0b974321b75d6ef64e7c3b118adb2801bf674282df5712cd2af390f79", "a3fc64fece6f3e1effab4045a9a54faa49a228f787025f0ecb761145755cb2d0",
        "3af178b78710adae9cc096188ed09c210078aaa7e965ef83d22a91f21fec4eb5", "688c743cde3987e299d2b028038ddc12dc02e7033c9d3c8f4d20edf9544232aa",
        "45e29166c6441f0fd27e3b85a205f1e102b025cc8e8ea158ab4885a22ed68905")])
    assert np.array_equal(op("to_classical", op("from_classical", vals)), vals)
    # Mgry.Ops  tests/mgry.cpp:78-120
    a = h("FFFFFFFFFFFFFFFFFFFFFF000000000000000000000000000000000000000004")
    b = h("FFFFFFFFFFFFFFFFFFFFFF000000000000000000000000000000000000000005")
    ma, mb = op("from_classical", a), op("from_classical", b)
    cl = lambda x: to_ints(op("to_classical", x))[0]
    assert cl(op("mod_add", ma, mb)) == 0xfffffffffffffffffffffe0000000000000000000000000000000001000003da
    assert cl(op("mod_sub", ma, mb)) == 0xfffffffffffffffffffffffffffffffffffffffffffffffffffffffefffffc2e
    assert cl(op("mod_sub", mb, ma)) == 1
    assert cl(op("mgry_pow", ma, e=K1 - 2)) == 0xDC1B98237FD316F9AEE7342E6DC7629A75A99A9E9EF591170282CE3E1D8E26ED   # also Mgry.Gfp inverse
    assert cl(op("mgry_pow", ma, e=2)) == 0xfffffffffffffdfffff85600000000000001000003d10001000007a9000eab68
    assert cl(op("mgry_pow", ma, e=0x00000000000F0000000000000000000000000000000000000000000000000001)) == \
        0xa51e978903ca7fcd788382ff283366ad7457d27c7aac417127a8723626773516
    assert cl(op("mgry_pow", ma, e=0)) == 1
    # Mgry.Gfp  tests/mgry.cpp:122-150: sqrt = pow((p+1)/4), opposite
    s = op("from_classical", h("b560fd7b259468b53c3a1623f35786a491fcb1fcdfbb0165da4dccce1f185b60"))
    assert cl(op("mgry_pow", s, e=(K1 + 1) // 4)) == 0xa59f1be7c1f892ff2adf14187e9cff7666112af579bc1a11b63e248098567e71
    assert not op("mod_add", s, op("opposite", s)).any()
    # Ops256.Mod  tests/ops.cpp:221-252 (plain modular ops on the secp256k1 prime)
    assert to_ints(op("mod_add", h("FFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEFFFFFC2E"), h("02")))[0] == 1
    x = h("fffffffffffffffffffffffffffffffffffffffffffffffffffffff000000000")
    y = h("ffeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeee")
    assert to_ints(op("mod_add", x, y))[0] == 0xffeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeeedfeeeef2bf
    assert to_ints(op("mod_sub", x, y))[0] == 0x0011111111111111111111111111111111111111111111111111110111111112
    assert to_ints(op("mod_shift_left_one", x))[0] == 0xffffffffffffffffffffffffffffffffffffffffffffffffffffffe1000003d1


def test_reference_kats_on_the_oracle(orc):
    _kats(lambda name, a, b=None, e=None: orc_op(orc, name, a, b, e))
    # Ops256.Binops  tests/ops.cpp:210-219: 256x256 -> 512
    prod = orc.mul512(to_words([2**256 - 1]), to_words([int("ee" * 32, 16)]))[0]
    assert sum(int(w) << (32 * i) for i, w in enumerate(prod)) == int("EE" * 31 + "ED" + "11" * 31 + "12", 16)


def test_generic_on_p256_equals_specialised_oracle(orc):
    a, b = _libs.field_elems(5, 512), _libs.field_elems(6, 512)
    P = _libs.P_INT
    assert np.array_equal(orc_op(orc, "mgry_mul", a, b, p=P), orc.mgry_mul(a, b))
    assert np.array_equal(orc_op(orc, "mgry_sqr", a, p=P), orc.mgry_sqr(a))
    assert np.array_equal(orc_op(orc, "mod_add", a, b, p=P), orc.mgry_add(a, b))
    assert np.array_equal(orc_op(orc, "mod_sub", a, b, p=P), orc.mgry_sub(a, b))
    assert np.array_equal(orc_op(orc, "opposite", a, p=P), orc.opposite(a))
    assert np.array_equal(orc_op(orc, "from_classical", a, p=P), orc.from_classical(a))


@pytest.mark.gpu
def test_engine_generic_prime(eng, orc):
    F = eng.GenericField(K1)
    a, b = _vals(4096, 11), _vals(4096, 12)
    for name in ("mod_add", "mod_sub", "mgry_mul"):
        assert np.array_equal(getattr(F, name)(a, b), orc_op(orc, name, a, b)), name
    for name in ("mod_shift_left_one", "mgry_sqr", "from_classical", "to_classical", "opposite"):
        assert np.array_equal(getattr(F, name)(a), orc_op(orc, name, a)), name
    q = _libs.quirk_stress(4000, 13)
    assert np.array_equal(F.mgry_sqr(q), orc_op(orc, "mgry_sqr", q))
    assert not np.array_equal(F.mgry_sqr(q, quirk=False), orc_op(orc, "mgry_sqr", q))
    for e in (K1 - 2, (K1 + 1) // 4, 0, 1, 5):
        assert np.array_equal(F.mgry_pow(a[:256], e), orc_op(orc, "mgry_pow", a[:256], e=e)), hex(e)
    # the reference's KATs straight on the engine
    op = lambda name, x, y=None, e=None: (F.mgry_pow(x, e) if name == "mgry_pow" else getattr(F, name)(x, y) if y is not None else getattr(F, name)(x))
    _kats(op)
    # 512-bit product and square (Ops256.Binops; square with the lost carry)
    assert np.array_equal(eng.mul512(a, b), orc.mul512(a, b))
    assert np.array_equal(eng.square512(q), orc.square512(q))
    prod = eng.mul512(to_words([2**256 - 1]), to_words([int("ee" * 32, 16)]))[0]
    assert sum(int(w) << (32 * i) for i, w in enumerate(prod)) == int("EE" * 31 + "ED" + "11" * 31 + "12", 16)
    # and the generic path agrees with the P-256 specialisation
    G = eng.GenericField(_libs.P_INT)
    c, d = _libs.field_elems(21, 2048), _libs.field_elems(22, 2048)
    assert np.array_equal(G.mgry_mul(c, d), eng.mgry_mul(c, d)) and np.array_equal(G.mgry_sqr(c), eng.mgry_sqr(c))
