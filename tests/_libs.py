"""Test-only loaders: the C oracle (oracle/libp256_oracle.so), the compiled
reference (oracle/_ref/libecsimd_ref.so, present only where /root/reference
was available at build time) and synthetic-input generators.

Nothing in the product package imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

P_INT = 2**256 - 2**224 + 2**192 + 2**96 - 1
N_INT = 0xFFFFFFFF00000000FFFFFFFFFFFFFFFFBCE6FAADA7179E84F3B9CAC2FC632551
R_INT = 2**256
B_INT = 0x5AC635D8AA3A93E7B3EBBD55769886BC651D06B0CC53B0F63BCE3C3E27D2604B
GX_INT = 0x6B17D1F2E12C4247F8BCE6E563A440F277037D812DEB33A0F4A13945D898C296
GY_INT = 0x4FE342E2FE1A7F9B8EE7EB4A7C0F9E162BCE33576B315ECECBB6406837BF51F5

_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")


def _build_oracle():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "libp256_oracle.so"], check=True)


def load_oracle():
    path = os.path.join(ORACLE_DIR, "libp256_oracle.so")
    src = os.path.join(ORACLE_DIR, "p256_oracle.c")
    if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
        _build_oracle()
    return C.CDLL(path)


def load_ref():
    """The compiled, unmodified reference; None where it could not be built."""
    path = os.path.join(ORACLE_DIR, "_ref", "libecsimd_ref.so")
    if not os.path.exists(path):
        return None
    try:
        with open("/proc/cpuinfo") as f:
            if " avx2" not in f.read():
                return None
    except OSError:
        return None
    return C.CDLL(path)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Lib:
    """Uniform numpy front-end over the oracle ("orc_") or reference ("ref_")
    array API.  Values are (n, 8) uint32 arrays, LS word first; Jacobian points
    (n, 24); affine points (n, 16)."""

    def __init__(self, lib, prefix, nt=1):
        self.lib, self.prefix, self.nt = lib, prefix, nt
        self.has_nt_raw = prefix == "orc_"

    def _f(self, name):
        return getattr(self.lib, self.prefix + name)

    def _call(self, name, outs, ins, n, threaded=True):
        args = [_ptr(o) for o in outs] + [_ptr(np.ascontiguousarray(i)) for i in ins] + [C.c_size_t(n)]
        if threaded:
            args.append(C.c_int(self.nt))
        f = self._f(name)
        f.restype = None
        f(*args)

    def _unary(self, name, a, wout=8):
        a = np.ascontiguousarray(a, dtype=np.uint32)
        n = a.shape[0]
        o = np.zeros((n, wout), np.uint32)
        self._call(name, [o], [a], n)
        return o

    def _binary(self, name, a, b, wout=8):
        a = np.ascontiguousarray(a, dtype=np.uint32)
        b = np.ascontiguousarray(b, dtype=np.uint32)
        n = a.shape[0]
        o = np.zeros((n, wout), np.uint32)
        self._call(name, [o], [a, b], n)
        return o

    def mgry_add(self, a, b): return self._binary("mgry_add", a, b)
    def mgry_sub(self, a, b): return self._binary("mgry_sub", a, b)
    def mgry_mul(self, a, b): return self._binary("mgry_mul", a, b)
    def mgry_shl1(self, a): return self._unary("mgry_shl1", a)
    def mgry_sqr(self, a): return self._unary("mgry_sqr", a)
    def opposite(self, a): return self._unary("opposite", a)
    def from_classical(self, a): return self._unary("from_classical", a)
    def to_classical(self, a): return self._unary("to_classical", a)
    def inverse(self, a): return self._unary("inverse", a)

    def mul512(self, a, b):
        a = np.ascontiguousarray(a, np.uint32); b = np.ascontiguousarray(b, np.uint32)
        o = np.zeros((a.shape[0], 16), np.uint32)
        self._call("mul512", [o], [a, b], a.shape[0], threaded=False)
        return o

    def square512(self, a):
        a = np.ascontiguousarray(a, np.uint32)
        o = np.zeros((a.shape[0], 16), np.uint32)
        self._call("square512", [o], [a], a.shape[0], threaded=False)
        return o

    def mgry_reduce(self, t):
        t = np.ascontiguousarray(t, np.uint32)
        o = np.zeros((t.shape[0], 8), np.uint32)
        self._call("mgry_reduce", [o], [t], t.shape[0], threaded=False)
        return o

    def _two_out(self, name, ins):
        ins = [np.ascontiguousarray(i, np.uint32) for i in ins]
        n = ins[0].shape[0]
        o1 = np.zeros((n, 24), np.uint32); o2 = np.zeros((n, 24), np.uint32)
        self._call(name, [o1, o2], ins, n)
        return o1, o2

    def dblu(self, P): return self._two_out("dblu", [P])          # (P', 2P)
    def trplu(self, P): return self._two_out("trplu", [P])        # (P', 3P)
    def zaddu(self, P, O): return self._two_out("zaddu", [P, O])  # (P', P+O)
    def zdau(self, P, Q): return self._two_out("zdau", [P, Q])    # (Q', 2P+Q)
    def add_z2_1(self, A, B): return self._binary("add_z2_1", A, B, 24)
    def scalar_mult(self, k, P): return self._binary("scalar_mult", k, P, 24)
    def from_affine(self, xy): return self._unary("from_affine", xy, 24)
    def to_affine(self, J): return self._unary("to_affine", J, 16)

    def constants(self):
        o = np.zeros(64, np.uint32)
        f = self._f("constants"); f.restype = None
        f(_ptr(o))
        return o.reshape(8, 8)


def oracle(nt=1):
    return Lib(load_oracle(), "orc_", nt)


def reference(nt=1):
    lib = load_ref()
    return None if lib is None else Lib(lib, "ref_", nt)


# ---- integer <-> word helpers --------------------------------------------------
def to_words(vals):
    """list of python ints -> (n, 8) uint32, LS word first"""
    out = np.zeros((len(vals), 8), np.uint32)
    for i, v in enumerate(vals):
        for j in range(8):
            out[i, j] = (v >> (32 * j)) & 0xFFFFFFFF
    return out


def to_ints(words):
    words = np.asarray(words, dtype=np.uint32).reshape(-1, 8)
    return [sum(int(w) << (32 * j) for j, w in enumerate(row)) for row in words]


def hexw(s):
    return to_words([int(s, 16)])[0]


# ---- synthetic inputs: counter-based splitmix64 -------------------
def splitmix64(x):
    """vectorised splitmix64 finaliser over uint64 arrays"""
    x = np.asarray(x, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = x + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def raw256(seed, n, start=0):
    """n raw 256-bit values: word(seed, 4*i + l) for limb l, as (n, 8) uint32"""
    idx = (np.arange(4 * n, dtype=np.uint64) + np.uint64(4 * start))
    with np.errstate(over="ignore"):
        w = splitmix64(np.uint64(seed) * np.uint64(0x100000001B3) + idx)
    return w.view(np.uint32).reshape(n, 8).copy()


def field_elems(seed, n, start=0):
    """canonical field elements: raw 256 bits, minus p once if >= p"""
    a = raw256(seed, n, start)
    ints = None
    # >= p is only possible when the top word is 0xffffffff: fix those rows
    rows = np.nonzero(a[:, 7] == 0xFFFFFFFF)[0]
    if len(rows):
        ints = to_ints(a[rows])
        a[rows] = to_words([v - P_INT if v >= P_INT else v for v in ints])
    return a


EDGE_FIELD = [0, 1, 2, P_INT - 1, P_INT - 2, R_INT % P_INT, (R_INT % P_INT) - 1, 2**255, 2**255 - 1,
              2**96 - 1, 2**96, 2**224, 2**192, 0xFFFFFFFF, 0x80000000, (2**256 - 1) % P_INT,
              0x8000000080000000800000008000000080000000800000008000000080000000,
              0x7FFFFFFF7FFFFFFF7FFFFFFF7FFFFFFF7FFFFFFF7FFFFFFF7FFFFFFF7FFFFFFF,
              0xFFFFFFFF00000000FFFFFFFF00000000FFFFFFFF00000000FFFFFFFF00000000 % P_INT,
              P_INT >> 1, (P_INT >> 1) + 1]
# squaring-quirk inputs: reference square() != a^2
QUIRK_FIELD = [0xA09D838E868B90F2B89CC416F270D3B8F0374E0A8728A79978B896A45AF4F8A8,
               0x196E98832350A697302E3812CF37CFDB65BD91769E220A413C2BC6519E220A41]
EDGE_SCALARS = [0, 1, 2, 3, 4, 5, N_INT - 1, N_INT, N_INT + 1, 2**256 - 1, 2**256 - 2, 2**255, 2**255 + 1,
                0x0BC1B1F28709DECB543D9677D2CC9942348F6B984DEFF409430740942FF38827,
                0x0A891CEC7F6B6F8E0F2B3F6CC9F5E51D0B1A7C2BF6B3F3E7C4D5A6B7C8D9BD80]


def quirk_stress(n, seed=1):
    """Field elements in which one digit pair (i, j) is chosen so that a_i*a_j lies just
    below 2^63 (high word 0x7fffffff): this is where the reference's square() can lose a
    carry (include/ecsimd/mul.h:192-206).  Roughly a quarter of them really do."""
    import random
    rnd = random.Random(seed)
    out = []
    while len(out) < n:
        d = [rnd.getrandbits(32) for _ in range(8)]
        i, j = sorted(rnd.sample(range(8), 2))
        ai = rnd.getrandbits(32) | 0x80000000
        aj = ((1 << 63) - 1 - rnd.getrandbits(20)) // ai
        if aj >= 1 << 32:
            continue
        if rnd.random() < 0.5:
            ai, aj = aj, ai
        d[i], d[j] = ai, aj
        if rnd.random() < 0.5:
            # favour large neighbours so that ret[]/prev are large too
            for k in range(8):
                if k not in (i, j) and rnd.random() < 0.5:
                    d[k] = 0xFFFFFFFF - rnd.getrandbits(8)
        v = sum(x << (32 * k) for k, x in enumerate(d))
        if v >= P_INT:
            continue
        out.append(v)
    return to_words(out)
