#!/usr/bin/env python3
"""Generator for ecsimd_b200/csrc/fp256_mul_gen.cuh.

Emits the P-256 Montgomery multiplication core as straight-line inline PTX and
checks the schedule offline with a bit-exact simulator (every PTX instruction
is modelled, every dropped carry is asserted to be impossible), so that a carry
bug is found here and not on the GPU.

Algorithm (what mgry_mul computes: include/ecsimd/mgry_ops.h:31-35 of the
reference = mul.h:150-158 followed by mgry_mul.h:84-121):
    t = (a*b + m*p) / 2^256,   m = -a*b*p^-1 mod 2^256,   r = t - p if t >= p
The reference does it as a 64-product schoolbook, then 8 reduction rounds of
9 multiplies each on AVX2 lanes.  Here it is re-designed for the sm_100a
integer pipes:

  * 32-bit limbs, coarsely-integrated operand scanning (CIOS): row i adds
    a*b_i and then m_i*p, and the lowest word (which is zero by construction)
    is dropped.
  * every 32x32+64 product is one IMAD.WIDE.U32(.X): ptxas fuses
    mad.lo.cc/madc.hi.cc on an aligned register pair, with the carry in a
    predicate.  To keep every product pair-aligned there are two accumulators,
    E (pairs at even absolute word positions) and O (odd positions); row i
    sends a_k*b_i to the accumulator whose parity is (i+k)&1.
  * p = 2^256 - 2^224 + 2^192 + 2^96 - 1, so -p^-1 = 1 mod 2^32 (m_i is just
    the current low word) and m_i*p = m_i*(p+1) - m_i: the "- m_i" cancels the
    low word exactly and (p+1) has only three non-zero digits {1@3, 1@6,
    0xffffffff@7}: three multiply-adds per row instead of nine.
  * the final merge E+O, and the conditional subtraction of p, are the only
    plain carry-chain adds.

Usage: python gen_fp256.py [--check-only]
"""
import os
import random
import sys

P = 2**256 - 2**224 + 2**192 + 2**96 - 1
M32 = 0xFFFFFFFF


class Emit:
    """Collects IR instructions grouped in asm statements; also simulates."""

    def __init__(self):
        self.stmts = []      # list of (list[str] ptx lines, outs, ins)
        self.cur = None
        self.ir = []         # flat IR for the simulator

    # --- statement grouping: one asm statement == one carry chain -------------
    def begin(self):
        assert self.cur is None
        self.cur = {"lines": [], "rw": [], "ro": [], "wo": []}

    def end(self):
        self.stmts.append(self.cur)
        self.ir.append(("endchain",))
        self.cur = None

    def _reg(self, name, mode):
        c = self.cur
        if isinstance(name, int):
            return str(name) if name < 2**31 else "0x%x" % name
        if mode == "w":
            if name in c["ro"]:
                c["ro"].remove(name); c["rw"].append(name)
            elif name not in c["rw"] and name not in c["wo"]:
                c["wo"].append(name)
        elif mode == "rw":
            if name in c["ro"]:
                c["ro"].remove(name)
            if name in c["wo"]:
                pass  # written earlier in this statement, then read: still write-only from outside
            elif name not in c["rw"]:
                c["rw"].append(name)
        else:  # read
            if name not in c["rw"] and name not in c["ro"] and name not in c["wo"]:
                c["ro"].append(name)
        return "%%{%s}" % name

    def line(self, fmt, *ops):
        # ops: (name, mode)
        txt = fmt.format(*[self._reg(n, m) for n, m in ops])
        self.cur["lines"].append(txt)

    # --- IR ops ----------------------------------------------------------------
    def mulw(self, lo, hi, x, y):
        """(hi:lo) = x*y   -- fresh pair, no carries"""
        self.line("mul.lo.u32 {0}, {1}, {2};", (lo, "w"), (x, "r"), (y, "r"))
        self.line("mul.hi.u32 {0}, {1}, {2};", (hi, "w"), (x, "r"), (y, "r"))
        self.ir.append(("mulw", lo, hi, x, y))

    def madw(self, lo, hi, x, y, cin, cout, fresh=(False, False)):
        """(hi:lo) (+)= x*y + cin ; carry-out kept iff cout; fresh = (lo is unset, hi is unset)"""
        if isinstance(fresh, bool):
            fresh = (fresh, fresh)
        a_lo = (0, "r") if fresh[0] else (lo, "rw")
        a_hi = (0, "r") if fresh[1] else (hi, "rw")
        self.line("mad%s.lo.cc.u32 {0}, {1}, {2}, {3};" % ("c" if cin else ""), (lo, "w" if fresh[0] else "rw"), (x, "r"), (y, "r"), a_lo)
        self.line("madc.hi%s.u32 {0}, {1}, {2}, {3};" % (".cc" if cout else ""), (hi, "w" if fresh[1] else "rw"), (x, "r"), (y, "r"), a_hi)
        self.ir.append(("madw", lo, hi, x, y, cin, cout, fresh))

    def addw(self, lo, hi, cin, cout):
        """(hi:lo) += cin  -- carry ripple through a pair"""
        assert cin
        self.line("addc.cc.u32 {0}, {0}, 0;", (lo, "rw"))
        self.line("addc%s.u32 {0}, {0}, 0;" % (".cc" if cout else ""), (hi, "rw"))
        self.ir.append(("addw", lo, hi, cin, cout))

    def add32(self, d, a, b, cin, cout, wrap_ok=False):
        op = "add" + ("c" if cin else "") + (".cc" if cout else "") + ".u32"
        self.line(op + " {0}, {1}, {2};", (d, "w" if (d != a and d != b) else "rw"), (a, "r" if a != d else "rw"), (b, "r" if b != d else "rw"))
        self.ir.append(("add32", d, a, b, cin, cout, wrap_ok))

    def sub32(self, d, a, b, cin, cout):
        op = "sub" + ("c" if cin else "") + (".cc" if cout else "") + ".u32"
        self.line(op + " {0}, {1}, {2};", (d, "w" if (d != a and d != b) else "rw"), (a, "r" if a != d else "rw"), (b, "r" if b != d else "rw"))
        self.ir.append(("sub32", d, a, b, cin, cout))

    def cap(self, d, fresh):
        """d (+)= carry"""
        if fresh:
            self.line("addc.u32 {0}, 0, 0;", (d, "w"))
        else:
            self.line("addc.u32 {0}, {0}, 0;", (d, "rw"))
        self.ir.append(("cap", d, fresh))

    def and32(self, d, a, imm):
        self.line("and.b32 {0}, {1}, %d;" % imm, (d, "w"), (a, "r"))
        self.ir.append(("and32", d, a, imm))

    # --- C++ text ----------------------------------------------------------------
    def cxx(self, indent="  "):
        out = []
        for s in self.stmts:
            names = s["rw"] + s["wo"] + s["ro"]
            idx = {n: i for i, n in enumerate(names)}
            body = []
            for ln in s["lines"]:
                t = ln
                for n in names:
                    t = t.replace("%%{%s}" % n, "%%%d" % idx[n])
                body.append(t)
            outs = ['"+r"(%s)' % n for n in s["rw"]] + ['"=r"(%s)' % n for n in s["wo"]]
            ins = ['"r"(%s)' % n for n in s["ro"]]
            # early-clobber for write-only outputs that are written before all inputs are read
            outs = [o.replace('"=r"', '"=&r"') for o in outs]
            out.append(indent + 'asm("' + (' "\n' + indent + '    "').join(body) + '"\n' + indent + "    : " + ", ".join(outs) +
                       "\n" + indent + "    : " + ", ".join(ins) + ");")
        return "\n".join(out)


def simulate(ir, env):
    """Bit-exact model of the emitted PTX; asserts on any lost carry."""
    cc = 0

    def val(x):
        return x if isinstance(x, int) else env[x]

    for ins in ir:
        k = ins[0]
        if k == "endchain":
            cc = None  # CC must not be consumed across statements
        elif k == "mulw":
            _, lo, hi, x, y = ins
            p = val(x) * val(y)
            env[lo], env[hi] = p & M32, p >> 32
        elif k == "madw":
            _, lo, hi, x, y, cin, cout, fresh = ins
            acc = ((0 if fresh[1] else env[hi]) << 32) | (0 if fresh[0] else env[lo])
            if cin:
                assert cc is not None
            t = acc + val(x) * val(y) + (cc if cin else 0)
            if not cout:
                assert t < 2**64, "lost carry in madw %s" % (ins,)
            env[lo], env[hi] = t & M32, (t >> 32) & M32
            cc = t >> 64
        elif k == "addw":
            _, lo, hi, cin, cout = ins
            assert cc is not None
            t = ((env[hi] << 32) | env[lo]) + cc
            if not cout:
                assert t < 2**64, "lost carry in addw"
            env[lo], env[hi] = t & M32, (t >> 32) & M32
            cc = t >> 64
        elif k == "add32":
            _, d, a, b, cin, cout, wrap_ok = ins
            if cin:
                assert cc is not None
            t = val(a) + val(b) + (cc if cin else 0)
            if not cout and not wrap_ok:
                assert t < 2**32, "lost carry in add32 %s" % (ins,)
            env[d] = t & M32
            cc = t >> 32
        elif k == "sub32":
            _, d, a, b, cin, cout = ins
            if cin:
                assert cc is not None
            t = val(a) - val(b) - (cc if cin else 0)
            env[d] = t & M32
            cc = 1 if t < 0 else 0   # CC.CF holds the borrow for sub
        elif k == "cap":
            _, d, fresh = ins
            assert cc is not None
            t = (0 if fresh else env[d]) + cc
            assert t < 2**32
            env[d] = t
            cc = None
        elif k == "and32":
            _, d, a, imm = ins
            env[d] = val(a) & imm
        else:
            raise ValueError(k)
    return env


def gen_mul(final="canonical"):
    """CIOS Montgomery product in E/O form.  Absolute word positions: E pair at
    even w is (e{w}, e{w+1}); O pair at odd w is (o{w}, o{w+1})."""
    g = Emit()
    touched = set()      # registers that hold a value (not fresh)

    def R(acc, w):
        return "%s%d" % (acc, w)

    def pair_fresh(acc, w):
        return R(acc, w) not in touched and R(acc, w + 1) not in touched

    def wfresh(acc, w):
        return (R(acc, w) not in touched, R(acc, w + 1) not in touched)

    def touch(acc, w):
        touched.add(R(acc, w)); touched.add(R(acc, w + 1))

    for i in range(8):
        Pn, Qn = ("e", "o") if i % 2 == 0 else ("o", "e")   # P has the parity of i
        b = "b%d" % i
        # ---- Q chain: [transfer] + odd-k products at words i+k ---------------
        g.begin()
        cin = False
        if i > 0:
            # word i lives in P (low word of pair (i,i+1)) and in Q (high word of pair (i-1,i))
            g.add32(R(Pn, i), R(Pn, i), R(Qn, i), False, True)
            cin = True
        for k in (1, 3, 5, 7):
            w = i + k
            last = k == 7
            if pair_fresh(Qn, w) and not cin:
                g.mulw(R(Qn, w), R(Qn, w + 1), "a%d" % k, b)
                cin = False
            else:
                fr = wfresh(Qn, w)
                g.madw(R(Qn, w), R(Qn, w + 1), "a%d" % k, b, cin, not last, fresh=fr)
                cin = not last
            touch(Qn, w)
        g.end()
        # ---- P chain: even-k products --------------------------------------------
        g.begin()
        cin = False
        allfresh = all(pair_fresh(Pn, i + k) for k in (0, 2, 4, 6))
        for k in (0, 2, 4, 6):
            w = i + k
            if allfresh:
                g.mulw(R(Pn, w), R(Pn, w + 1), "a%d" % k, b)
            else:
                fr = wfresh(Pn, w)
                g.madw(R(Pn, w), R(Pn, w + 1), "a%d" % k, b, cin, True, fresh=fr)
                cin = True
            touch(Pn, w)
        if not allfresh:
            top = R(Pn, i + 8)
            g.cap(top, top not in touched)
            touched.add(top)
        g.end()
        m = R(Pn, i)
        # ---- reduction, Q side: m@i+3, ripple@i+5, m*0xffffffff@i+7 -----------------
        g.begin()
        g.madw(R(Qn, i + 3), R(Qn, i + 4), m, 1, False, True, fresh=wfresh(Qn, i + 3))
        g.addw(R(Qn, i + 5), R(Qn, i + 6), True, True)
        g.madw(R(Qn, i + 7), R(Qn, i + 8), m, 0xFFFFFFFF, True, True, fresh=wfresh(Qn, i + 7))
        top = R(Qn, i + 9)
        g.cap(top, top not in touched)
        touched.add(top)
        g.end()
        # ---- reduction, P side: m@i+6 ------------------------------------------------
        g.begin()
        g.madw(R(Pn, i + 6), R(Pn, i + 7), m, 1, False, True, fresh=wfresh(Pn, i + 6))
        top = R(Pn, i + 8)
        g.cap(top, top not in touched)
        touched.add(top)
        g.end()
    # ---- merge t = E[8..16] + O[8..16] ------------------------------------------------
    g.begin()
    for w in range(8, 17):
        ew, ow = R("e", w), R("o", w)
        ev = ew if ew in touched else 0
        ov = ow if ow in touched else 0
        g.add32("t%d" % (w - 8), ev, ov, w > 8, w < 16)
    g.end()
    if final == "canonical":
        # d = t - p (9 words); mask = all-ones iff t < p; r = d + (p & mask)
        pw = [(P >> (32 * k)) & M32 for k in range(8)]
        g.begin()
        for k in range(8):
            g.sub32("d%d" % k, "t%d" % k, pw[k], k > 0, True)
        g.sub32("mk", "t8", 0, True, False)
        g.end()
        g.begin()
        g.and32("m1", "mk", 1)
        g.end()
        g.begin()
        addp = ["mk", "mk", "mk", 0, 0, 0, "m1", "mk"]
        for k in range(8):
            g.add32("r%d" % k, "d%d" % k, addp[k], k > 0, k < 7, wrap_ok=True)  # mod 2^256
        g.end()
    return g, touched


def check(g, ntests=3000, seed=1):
    rnd = random.Random(seed)
    specials = [0, 1, P - 1, P - 2, 2**256 - 1, 2**255, 2**224 - 1, M32, (2**256 - 1) ^ (M32 << 96),
                int("ffffffff" * 8, 16), int("80000000" * 8, 16), int("7fffffff" * 8, 16),
                int("ffffffff00000000" * 4, 16), int("00000000ffffffff" * 4, 16), 2**96 - 1, 2**192, P >> 1]
    cases = [(x, y) for x in specials for y in specials]
    for _ in range(ntests):
        bits = rnd.choice([256, 256, 256, 255, 200, 64])
        x = rnd.getrandbits(bits); y = rnd.getrandbits(rnd.choice([256, 256, 224, 32]))
        if rnd.random() < 0.2:
            # words of all-ones / all-zeros patterns stress the carry paths
            x = int("".join(rnd.choice(["ffffffff", "00000000", "%08x" % rnd.getrandbits(32)]) for _ in range(8)), 16)
            y = int("".join(rnd.choice(["ffffffff", "00000000", "%08x" % rnd.getrandbits(32)]) for _ in range(8)), 16)
        cases.append((x, y))
    Rinv = pow(2**256, -1, P)
    for x, y in cases:
        env = {}
        for k in range(8):
            env["a%d" % k] = (x >> (32 * k)) & M32
            env["b%d" % k] = (y >> (32 * k)) & M32
        simulate(g.ir, env)
        r = sum(env["r%d" % k] << (32 * k) for k in range(8))
        T = x * y
        m = (-T * pow(P, -1, 2**256)) % 2**256
        t = (T + m * P) >> 256
        want = t - P if t >= P else t
        want &= 2**256 - 1
        assert r == want, "mismatch for %x * %x: got %x want %x" % (x, y, r, want)
        if x < P and y < P:
            assert r == (x * y * Rinv) % P
    return len(cases)


HEADER = '''// GENERATED by gen_fp256.py -- do not edit by hand; edit the generator.
//
// P-256 Montgomery multiplication core for sm_100a: r = a*b*2^-256 mod p,
// canonical in [0,p) (for ANY 256-bit a,b: the exact quotient t=(ab+mp)/2^256
// with one conditional subtraction, which is what the reference's
// mgry_mul computes: include/ecsimd/mgry_ops.h:31-35, mul.h:150-158,
// mgry_mul.h:84-121).
//
// Schedule: CIOS on 32-bit limbs with two pair-aligned accumulators (E: even
// word positions, O: odd) so that every 32x32+64 multiply-add is a single
// IMAD.WIDE.U32(.X) with the carry in a predicate; the reduction by
// p+1 = {1@3, 1@6, 0xffffffff@7} needs three multiply-adds per row and no
// multiplication by m' (m' = 1).  See gen_fp256.py for the derivation and the
// offline carry-bound simulation.
#pragma once
#include <cstdint>

namespace ecb200 {

'''


def emit_header(path):
    g, touched = gen_mul()
    n = check(g)
    regs = sorted(touched, key=lambda s: (s[0], int(s[1:])))
    txt = HEADER
    txt += "__device__ __forceinline__ void fp_mul_words(\n"
    txt += "    uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t& r4, uint32_t& r5, uint32_t& r6, uint32_t& r7,\n"
    txt += "    uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t a4, uint32_t a5, uint32_t a6, uint32_t a7,\n"
    txt += "    uint32_t b0, uint32_t b1, uint32_t b2, uint32_t b3, uint32_t b4, uint32_t b5, uint32_t b6, uint32_t b7) {\n"
    txt += "  uint32_t " + ", ".join(regs) + ";\n"
    txt += "  uint32_t " + ", ".join("t%d" % k for k in range(9)) + ";\n"
    txt += "  uint32_t " + ", ".join("d%d" % k for k in range(8)) + ", mk, m1;\n"
    txt += g.cxx() + "\n}\n\n}  // namespace ecb200\n"
    with open(path, "w") as f:
        f.write(txt)
    return n, g


if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    if "--check-only" in sys.argv:
        g, _ = gen_mul()
        print("simulated cases ok:", check(g, 20000))
    else:
        n, g = emit_header(os.path.join(here, "fp256_mul_gen.cuh"))
        nimad = sum(1 for i in g.ir if i[0] in ("mulw", "madw"))
        print("wrote fp256_mul_gen.cuh; simulated %d cases ok; wide multiply-adds: %d" % (n, nimad))
