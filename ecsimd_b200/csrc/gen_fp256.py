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

    def raw(self, text, ir, regs=()):
        """a plain C++ statement between two asm statements (modelled by the IR tuple `ir`)"""
        assert self.cur is None
        self.stmts.append({"raw": text, "lines": [], "rw": [], "ro": [], "wo": list(regs)})
        self.ir.append(ir)

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
        # one 64-bit multiply, so that ptxas cannot split it into IMAD + IMAD.HI (5 clk) when the two
        # halves have different consumers
        self.line("{{ .reg .b64 w; mul.wide.u32 w, {2}, {3}; mov.b64 {{{0}, {1}}}, w; }}", (lo, "w"), (hi, "w"), (x, "r"), (y, "r"))
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

    def capb(self, d):
        """d = 0 - 0 - borrow  (0 or 0xffffffff)"""
        self.line("subc.u32 {0}, 0, 0;", (d, "w"))
        self.ir.append(("capb", d))

    def setc(self, c):
        """carry flag := (c != 0) for c in {0,1}"""
        self.line("add.cc.u32 {0}, {1}, 0xffffffff;", ("scratch", "w"), (c, "r"))
        self.ir.append(("setc", c))

    def setb(self, b):
        """borrow flag := (b != 0) for b in {0, 0xffffffff}"""
        self.line("sub.cc.u32 {0}, 0, {1};", ("scratch", "w"), (b, "r"))
        self.ir.append(("setb", b))

    def shl32(self, d, a, n):
        self.line("shl.b32 {0}, {1}, %d;" % n, (d, "w"), (a, "r"))
        self.ir.append(("shl32", d, a, n))

    def shf_l(self, d, lo, hi, n):
        """d = (hi:lo) << n, upper word (funnel shift)"""
        self.line("shf.l.clamp.b32 {0}, {1}, {2}, %d;" % n, (d, "w"), (lo, "r"), (hi, "r"))
        self.ir.append(("shf_l", d, lo, hi, n))

    def and32(self, d, a, imm):
        self.line("and.b32 {0}, {1}, %d;" % imm, (d, "w"), (a, "r"))
        self.ir.append(("and32", d, a, imm))

    # --- C++ text ----------------------------------------------------------------
    def cxx(self, indent="  "):
        out = []
        for s in self.stmts:
            if "raw" in s:
                out.append(indent + s["raw"])
                continue
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
            if not cout:
                assert t >= 0, "lost borrow in sub32 %s" % (ins,)
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
        elif k == "capb":
            assert cc is not None
            env[ins[1]] = M32 if cc else 0
            cc = None
        elif k == "setc":
            assert val(ins[1]) in (0, 1)
            cc = val(ins[1])
        elif k == "setb":
            assert val(ins[1]) in (0, M32)
            cc = 1 if val(ins[1]) else 0
        elif k == "copy":
            env[ins[1]] = val(ins[2])
        elif k == "max3s":
            def sgn(v):
                return v - 2**32 if v >= 2**31 else v
            env[ins[1]] = max(sgn(val(x)) for x in ins[2:]) & M32
        elif k == "shl32":
            _, d, a, n = ins
            env[d] = (val(a) << n) & M32
        elif k == "shf_l":
            _, d, lo, hi, n = ins
            env[d] = (((val(hi) << 32) | val(lo)) << n >> 32) & M32
        else:
            raise ValueError(k)
    return env


# ---------------------------------------------------------------------------------------------
# Schedules.  Absolute word positions: the E accumulator holds pairs at even positions
# (e{w}, e{w+1}), the O accumulator pairs at odd positions (o{w}, o{w+1}).  A "chain" is a
# list of (position, x, y) products at strictly ascending consecutive pairs of ONE accumulator;
# the carry travels from one multiply-add to the next in a predicate.  A chain may end without
# capturing its carry only where the carry cannot exist: on a pair nothing has touched yet
# (a*b + carry < 2^64) or on the top pair of the accumulator (the whole sum is bounded).
# Every 32x32+64 step is one IMAD.WIDE.U32(.X): 4.1 clk of the FMA-heavy pipe per warp on B200
# (measured, tools/microbench_mix.py), so the schedules below use exactly 64 (mul) / 36 (sqr)
# of them and keep everything else on plain IADD3 carry chains.
# Order matters: a chain that ends on an already used pair hands its carry-out to the NEXT pair
# as a one-instruction "deposit" (addc into a still untouched register), and it is scheduled
# before any product touches that next pair; the chain that later ends there still absorbs its
# own carry, because a*b + deposit + carry < 2^64.  So the 64 products need 7 deposits and no
# extra carry-propagation chain.
MUL_E_CHAINS = [
    [(0, 0, 0), (2, 0, 2), (4, 0, 4), (6, 0, 6)],                                  # ends fresh on pair 6
    [(2, 2, 0), (4, 2, 2), (6, 2, 4), (8, 2, 6)],                                  # ends fresh on 8
    [(8, 7, 1)],                                                                   # used pair 8 -> deposit on 10
    [(6, 5, 1), (8, 5, 3), (10, 7, 3)],                                            # ends on 10 (deposit only)
    [(2, 1, 1), (4, 1, 3), (6, 1, 5), (8, 1, 7), (10, 3, 7)],                      # used 10 -> deposit on 12
    [(6, 6, 0), (8, 6, 2), (10, 6, 4), (12, 7, 5)],                                # ends on 12 (deposit only)
    [(4, 4, 0), (6, 4, 2), (8, 4, 4), (10, 4, 6), (12, 5, 7)],                     # used 12 -> deposit on 14
    [(4, 3, 1), (6, 3, 3), (8, 3, 5), (10, 5, 5), (12, 6, 6), (14, 7, 7)],         # ends on 14 (top)
]
MUL_O_CHAINS = [
    [(1, 0, 1), (3, 0, 3), (5, 0, 5), (7, 0, 7)],                                  # ends fresh on 7
    [(7, 7, 0)],                                                                   # used 7 -> deposit on 9
    [(7, 6, 1), (9, 6, 3)],                                                        # ends on 9 (deposit only)
    [(1, 1, 0), (3, 1, 2), (5, 1, 4), (7, 1, 6), (9, 3, 6)],                       # used 9 -> deposit on 11
    [(5, 5, 0), (7, 5, 2), (9, 7, 2), (11, 7, 4)],                                 # ends on 11 (deposit only)
    [(3, 2, 1), (5, 2, 3), (7, 2, 5), (9, 2, 7), (11, 4, 7)],                      # used 11 -> deposit on 13
    [(5, 4, 1), (7, 4, 3), (9, 4, 5), (11, 6, 5), (13, 6, 7)],                     # ends on 13 (deposit only)
    [(3, 3, 0), (5, 3, 2), (7, 3, 4), (9, 5, 4), (11, 5, 6), (13, 7, 6)],          # used 13 -> deposit on word 15
]
# cross products a_i*a_j, i<j, of the squaring.  Seven chains of four that each END on a pair nothing has touched yet
# (no deposits), ordered so that the first product to land on every pair -- it has an RZ addend -- is one whose two
# limbs are in the same class of S = {0, 1, 2, 4}, T = {3, 5, 6, 7}: the twelve products inside S and inside T have
# twelve different sums i+j, so each can be the first on its pair, and the 15 products that accumulate onto a
# register pair all pair a limb of S with a limb of T.  An IMAD.WIDE with a register addend takes an extra issue
# clock when its two multiplicands sit in the same register bank (profiles/r2_bank_conflicts.md); with this
# schedule an allocation "S in one bank, T in the other" (csrc/sass_recolor.py) leaves no such multiply in a squaring
# (any schedule whose accumulating products contain a triangle keeps at least one per triangle: the round-1 one
# had all 15 pairs among the limbs 1..6).
SQR_E_CHAINS = [
    [(2, 0, 2), (4, 0, 4), (6, 2, 4), (8, 3, 5)],          # all four first on their pair: exact products
    [(4, 1, 3), (6, 0, 6), (8, 1, 7), (10, 3, 7)],
    [(6, 1, 5), (8, 2, 6), (10, 4, 6), (12, 5, 7)],
]
SQR_O_CHAINS = [
    [(1, 0, 1), (3, 1, 2), (5, 1, 4), (7, 0, 7)],          # all four first on their pair: exact products
    [(3, 0, 3), (5, 0, 5), (7, 1, 6), (9, 3, 6)],
    [(5, 2, 3), (7, 2, 5), (9, 2, 7), (11, 5, 6)],
    [(7, 3, 4), (9, 4, 5), (11, 4, 7), (13, 6, 7)],
]
SQR_BANK_CLASSES = ({0, 1, 2, 4}, {3, 5, 6, 7})


def _check_sqr_banks():
    """every product that accumulates onto a register pair (i.e. is not the first to touch it) pairs S with T"""
    S, T = SQR_BANK_CLASSES
    for chains in (SQR_E_CHAINS, SQR_O_CHAINS):
        touched = set()
        for ch in chains:
            for (w, i, j) in ch:
                if w in touched:
                    assert (i in S) != (j in S), "accumulating product a%d*a%d is inside one bank class" % (i, j)
                touched.add(w)


def _check_cover(chains_e, chains_o, square):
    seen = set()
    for acc, chains in (("e", chains_e), ("o", chains_o)):
        for ch in chains:
            prev = None
            for (w, i, j) in ch:
                assert w == i + j and (w % 2 == 0) == (acc == "e")
                assert prev is None or w == prev + 2
                assert (i, j) not in seen
                seen.add((i, j))
                prev = w
    want = {(i, j) for i in range(8) for j in range(8) if (i < j if square else True)}
    assert seen == want, (sorted(want - seen), sorted(seen - want))


def emit_products(g, chains_e, chains_o, xa, xb, tops, fresh_hook=None):
    """Emit the product chains.  Returns the set of registers that hold a value afterwards.
    fresh_hook(list of (i, j, hi register)) is called after every chain whose products all landed
    on untouched pairs: there the registers hold the exact products a_i*b_j."""
    touched = set()       # registers holding a value
    deposit_only = set()  # ... that so far hold nothing but a deposited carry (0/1)

    def R(acc, w):
        return "%s%d" % (acc, w)
    for acc, chains in (("e", chains_e), ("o", chains_o)):
        top = tops[acc]
        for ch in chains:
            allfresh = all(R(acc, w) not in touched and R(acc, w + 1) not in touched for (w, _, _) in ch)
            g.begin()
            cin = False
            for n, (w, i, j) in enumerate(ch):
                lo, hi = R(acc, w), R(acc, w + 1)
                last = n == len(ch) - 1
                fresh = (lo not in touched, hi not in touched)
                if allfresh and not (n == 0 and w == 0 and not MULW_WORD0):
                    g.mulw(lo, hi, xa % i, xb % j)
                elif allfresh:
                    # word 0 of the product is m_0, read four more times by the reduction: written as a
                    # carry-setting multiply-add so that ptxas keeps it instead of recomputing it
                    g.madw(lo, hi, xa % i, xb % j, False, True, fresh=(True, True))
                else:
                    # the carry out of the last step is dropped only where it cannot exist: on a pair that
                    # holds at most a deposited carry, or on the top pair of the accumulator
                    nearly_fresh = fresh[1] and (fresh[0] or lo in deposit_only)
                    safe_end = last and (nearly_fresh or w == top)
                    g.madw(lo, hi, xa % i, xb % j, cin, not safe_end, fresh=fresh)
                    cin = True
                    if last and not safe_end:
                        dep = R(acc, w + 2)
                        assert dep not in touched, "deposit target %s already in use: reorder the chains" % dep
                        g.cap(dep, True)
                        touched.add(dep)
                        deposit_only.add(dep)
                touched.add(lo); touched.add(hi)
                deposit_only.discard(lo)
            g.end()
            if allfresh and fresh_hook is not None:
                fresh_hook([(i, j, R(acc, w + 1)) for (w, i, j) in ch])
    return touched


PIN_M0 = True
MULW_WORD0 = False  # True: a_0*b_0 as mul.wide too (the m_0 pin then costs one ALU add instead of an IMAD): no gain in the ladder, -4.7 % on the register-resident multiply chain (ptxas schedules it worse)


def emit_reduction(g, T):
    """Montgomery reduction of the 16-word T (register names T[0..15]) for p256, on plain carry
    chains.  m = -T/p mod 2^256 is found word by word (m' = 1) from
        m = low256(T + (m<<96) + (m<<192) - (m<<224))
    and the quotient is t = (T + (m<<96) + (m<<192) - (m<<224) + (m<<256)) >> 256, nine words
    h0..h8 (h8 is 0 or 1); the conditional subtraction of p is left to the caller."""
    if PIN_M0:
        # ptxas otherwise re-computes the low word of a0*b0 with an extra IMAD at each of its four uses
        # (it is cheaper than a register in its cost model); adding a zero it cannot see through pins it.
        g.raw("m0 = %s + ecb200_opaque_zero();" % T[0], ("copy", "m0", T[0]), regs=["m0"])
        T = ["m0"] + list(T[1:])
    m = ["%s" % T[0], "%s" % T[1], "%s" % T[2], "m3", "m4", "m5", "m6", "m7"]
    # low words of +(m<<192) and -(m<<224): they need only m0, m1
    g.begin()
    g.add32("x6", T[6], m[0], False, True)
    g.add32("x7", T[7], m[1], True, True)
    g.cap("cB", True)
    g.end()
    g.begin()
    g.sub32("y7", "x7", m[0], False, True)
    g.capb("bC")
    g.end()
    # +(m<<96): words 3..7 give m3..m7, and the same chain runs on through the high half
    g.begin()
    g.add32("m3", T[3], m[0], False, True)
    g.add32("m4", T[4], m[1], True, True)
    g.add32("m5", T[5], m[2], True, True)
    g.add32("m6", "x6", "m3", True, True)
    g.add32("m7", "y7", "m4", True, True)
    for k in range(8):
        src = m[5 + k] if 5 + k < 8 else 0
        g.add32("h%d" % k, T[8 + k], src, True, True)
    g.cap("h8", True)
    g.end()
    # +(m<<192), high half: m2..m7 at h0..h5, carry-in = cB
    g.begin()
    g.setc("cB")
    for k in range(8):
        src = m[2 + k] if 2 + k < 8 else 0
        g.add32("h%d" % k, "h%d" % k, src, True, True)
    g.cap("h8", False)
    g.end()
    # +(m<<256): m0..m7 at h0..h7
    g.begin()
    for k in range(8):
        g.add32("h%d" % k, "h%d" % k, m[k], k > 0, True)
    g.cap("h8", False)
    g.end()
    # -(m<<224), high half: m1..m7 at h0..h6, borrow-in = bC
    g.begin()
    g.setb("bC")
    for k in range(8):
        src = m[1 + k] if 1 + k < 8 else 0
        g.sub32("h%d" % k, "h%d" % k, src, True, True)
    g.sub32("h8", "h8", 0, True, False)
    g.end()


def gen_mul():
    _check_cover(MUL_E_CHAINS, MUL_O_CHAINS, False)
    g = Emit()
    touched = emit_products(g, MUL_E_CHAINS, MUL_O_CHAINS, "a%d", "b%d", {"e": 14, "o": None})
    # merge T = E + O (word 0 is e0 itself)
    g.begin()
    for w in range(1, 16):
        ev = "e%d" % w if "e%d" % w in touched else 0
        ov = "o%d" % w if "o%d" % w in touched else 0
        g.add32("t%d" % w, ev, ov, w > 1, w < 15)
    g.end()
    emit_reduction(g, ["e0"] + ["t%d" % w for w in range(1, 16)])
    return g


def gen_mul512():
    """product + merge only: T = a*b as 16 words t0..t15 (used by the generic-prime path)"""
    g = Emit()
    touched = emit_products(g, MUL_E_CHAINS, MUL_O_CHAINS, "a%d", "b%d", {"e": 14, "o": None})
    g.begin()
    g.add32("t0", "e0", 0, False, False)
    g.end()
    g.begin()
    for w in range(1, 16):
        ev = "e%d" % w if "e%d" % w in touched else 0
        ov = "o%d" % w if "o%d" % w in touched else 0
        g.add32("t%d" % w, ev, ov, w > 1, w < 15)
    g.end()
    return g


def gen_mul_acc():
    """t = (a*b + c + m*p) / 2^256 for a 512-bit addend c (16 words c0..c15, any value): the product
    chains as in gen_mul, the merge takes c as a second chain, the 17th word joins h8 (0..2).
    Used to share one unreduced product between two results (point.cuh: Y3 and Y2n)."""
    g = Emit()
    touched = emit_products(g, MUL_E_CHAINS, MUL_O_CHAINS, "a%d", "b%d", {"e": 14, "o": None})
    g.begin()
    for w in range(1, 16):
        ev = "e%d" % w if "e%d" % w in touched else 0
        ov = "o%d" % w if "o%d" % w in touched else 0
        g.add32("u%d" % w, ev, ov, w > 1, w < 15)
    g.end()
    g.begin()
    g.add32("t0", "e0", "c0", False, True)
    for w in range(1, 16):
        g.add32("t%d" % w, "u%d" % w, "c%d" % w, True, True)
    g.cap("t16", True)
    g.end()
    emit_reduction(g, ["t%d" % w for w in range(16)])
    g.begin()
    g.add32("h8", "h8", "t16", False, False)
    g.end()
    return g


def gen_sqr():
    _check_cover(SQR_E_CHAINS, SQR_O_CHAINS, True)
    _check_sqr_banks()
    g = Emit()
    g.exact_pairs = []   # cross products whose high word is observed exactly (see fp_sqr_quirk_filter)

    def hook(items):
        # qx = signed max of (its value on entry and) the exact high words hi32(a_i*a_j): INT_MAX <=>
        # some pair is in the reference's lost-carry set (fp256.cuh).  qx is in/out so that a group
        # of squarings shares one accumulator; two new values per VIMNMX3.
        regs = [r for (_, _, r) in items]
        g.exact_pairs.extend((i, j) for (i, j, _) in items)
        while regs:
            take = regs[:2]
            regs = regs[2:]
            args = ["qx"] + take
            while len(args) < 3:
                args.append(args[-1])
            g.raw("qx = (uint32_t)__vimax3_s32((int)%s, (int)%s, (int)%s);" % tuple(args), ("max3s", "qx") + tuple(args))
    touched = emit_products(g, SQR_E_CHAINS, SQR_O_CHAINS, "a%d", "a%d", {"e": None, "o": None}, fresh_hook=hook)
    # cross sum S = E + O: words 1..14, carry into word 15
    g.begin()
    for w in range(1, 15):
        ev = "e%d" % w if "e%d" % w in touched else 0
        ov = "o%d" % w if "o%d" % w in touched else 0
        if ev == 0 and ov == 0:
            continue
        g.add32("s%d" % w, ev, ov, w > 1, True)
    g.cap("s15", True)
    g.end()
    # doubled: d = 2*S (words 1..15; bit 0 of word 1 is zero)
    g.begin()
    g.shl32("d1", "s1", 1)
    for w in range(2, 16):
        g.shf_l("d%d" % w, "s%d" % (w - 1), "s%d" % w, 1)
    g.end()
    # T = d + sum a_i^2 2^(64 i): one chain of eight multiply-adds on the pairs (2i, 2i+1)
    # a_0^2 as a fresh 64-bit product whose high word is then added to d1: as a multiply-add onto (nothing, d1) ptxas
    # splits it into IMAD + IMAD.HI (2 + 5.1 issue clocks instead of 4.1 + 1.2)
    g.begin()
    g.mulw("t0", "x1", "a0", "a0")
    g.add32("d1", "d1", "x1", False, True)
    for i in range(1, 8):
        g.madw("d%d" % (2 * i), "d%d" % (2 * i + 1), "a%d" % i, "a%d" % i, True, i < 7)
    g.end()
    T = ["t0"] + ["d%d" % w for w in range(1, 16)]
    emit_reduction(g, T)
    return g


def _cases(ntests, seed, unary):
    rnd = random.Random(seed)
    specials = [0, 1, P - 1, P - 2, 2**256 - 1, 2**255, 2**224 - 1, M32, (2**256 - 1) ^ (M32 << 96),
                int("ffffffff" * 8, 16), int("80000000" * 8, 16), int("7fffffff" * 8, 16),
                int("ffffffff00000000" * 4, 16), int("00000000ffffffff" * 4, 16), 2**96 - 1, 2**192, P >> 1]
    cases = [(x, x) for x in specials] if unary else [(x, y) for x in specials for y in specials]

    def pattern():
        return int("".join(rnd.choice(["ffffffff", "00000000", "%08x" % rnd.getrandbits(32)]) for _ in range(8)), 16)
    for _ in range(ntests):
        x = rnd.getrandbits(rnd.choice([256, 256, 256, 255, 200, 64]))
        y = rnd.getrandbits(rnd.choice([256, 256, 224, 32]))
        if rnd.random() < 0.25:      # words of all-ones / all-zeros stress the carry paths
            x, y = pattern(), pattern()
        cases.append((x, x) if unary else (x, y))
    return cases


def check(g, unary=False, ntests=3000, seed=1):
    """Simulate the emitted PTX on edge and random operands (ANY 256-bit pattern, not only
    canonical ones) and compare the nine-word quotient with the exact value."""
    pinv = pow(P, -1, 2**256)
    cases = _cases(ntests, seed, unary)
    for x, y in cases:
        env = {}
        for k in range(8):
            env["a%d" % k] = (x >> (32 * k)) & M32
            env["b%d" % k] = (y >> (32 * k)) & M32
        qx_in = env["qx"] = (0x80000000, 0x12345678, 0x7ffffffe)[(x ^ y) % 3]   # qx is in/out in the squaring
        simulate(g.ir, env)
        got = sum(env["h%d" % k] << (32 * k) for k in range(9))
        if unary and getattr(g, "exact_pairs", None):
            def sgn(v):
                return v - 2**32 if v >= 2**31 else v
            want_qx = max([sgn(qx_in)] + [sgn((env["a%d" % i] * env["a%d" % j]) >> 32) for (i, j) in g.exact_pairs]) & M32
            assert env["qx"] == want_qx, "qx mismatch"
        T = x * y
        m = (-T * pinv) % 2**256
        want = (T + m * P) >> 256
        assert got == want, "mismatch for %x * %x: got %x want %x" % (x, y, got, want)
    return len(cases)


HEADER = """// GENERATED by gen_fp256.py -- do not edit by hand; edit the generator.
//
// P-256 Montgomery multiplication / squaring cores for sm_100a.  Both return the exact
// nine-word quotient t = (T + m*p) / 2^256 with m = -T/p mod 2^256 (t < 2^256 + p), for T = a*b
// resp. T = a*a and ANY 256-bit operands; the caller (fp256.cuh) subtracts p once if t >= p,
// which is what the reference's mgry_mul / mgry_reduce compute
// (include/ecsimd/mgry_ops.h:31-35, mul.h:150-158, mgry_mul.h:84-121).
//
// Design (see gen_fp256.py for the schedules and the offline carry-bound simulation):
//  * 32-bit limbs; every 32x32+64 product is ONE IMAD.WIDE.U32(.X) (ptxas fuses
//    mad.lo.cc/madc.hi.cc on an aligned register pair, carry in a predicate).  To keep every
//    product pair-aligned there are two accumulators: E (pairs at even word positions) takes
//    a_i*b_j with i+j even, O the others.  Products are strung into chains of ascending pairs
//    that end only where no carry can exist, so the multiplier costs exactly 64 (squaring: 36)
//    IMAD.WIDE and 6 (0) captured carries.
//  * p = 2^256 - 2^224 + 2^192 + 2^96 - 1 gives -1/p = 1 mod 2^32 and m*p = (m<<256) - (m<<224)
//    + (m<<192) + (m<<96) - m: the reduction is four shifted multi-word add/sub chains on the
//    ALU pipe and needs no multiplication at all.
#pragma once
#include <cstdint>

namespace ecb200 {

// A zero the compiler cannot see through (a __constant__ word nobody writes): used to pin values
// that ptxas would otherwise rematerialise with extra multiplies.
__constant__ uint32_t g_ecb200_opaque_zero;
__device__ __forceinline__ uint32_t ecb200_opaque_zero() { return g_ecb200_opaque_zero; }

"""


def _emit_fn(name, g, nin, outs=None, extra_in=()):
    regs = set()
    for s in g.stmts:
        regs.update(s["rw"]); regs.update(s["wo"]); regs.update(s["ro"])
    outs = outs or ["h%d" % k for k in range(9)]
    params = outs + ["a%d" % k for k in range(8)] + (["b%d" % k for k in range(8)] if nin == 2 else []) + list(extra_in)
    local = sorted(regs - set(params), key=lambda x: (x.rstrip("0123456789"), int("0" + "".join(ch for ch in x if ch.isdigit()))))
    txt = "__device__ __forceinline__ void %s(\n" % name
    txt += "    " + ", ".join("uint32_t& %s" % o for o in outs) + ",\n"
    txt += "    " + ", ".join("uint32_t a%d" % k for k in range(8))
    if nin == 2:
        txt += ",\n    " + ", ".join("uint32_t b%d" % k for k in range(8))
    if extra_in:
        txt += ",\n    " + ", ".join("uint32_t %s" % n for n in extra_in)
    txt += ") {\n"
    txt += "  uint32_t " + ", ".join(local) + ";\n"
    txt += g.cxx() + "\n}\n\n"
    return txt


def emit_header(path):
    gm, gs = gen_mul(), gen_sqr()
    nm = check(gm, unary=False)
    ns = check(gs, unary=True)
    g5 = gen_mul512()
    ga = gen_mul_acc()
    pinv = pow(P, -1, 2**256)
    for x, y in _cases(1500, 7, False):
        for c in (0, x * y ^ (x << 200), 2**512 - 1, (P - 1) * (P - 1), (x * x) % 2**512):
            env = {}
            for k in range(8):
                env["a%d" % k] = (x >> (32 * k)) & M32
                env["b%d" % k] = (y >> (32 * k)) & M32
            for k in range(16):
                env["c%d" % k] = (c >> (32 * k)) & M32
            simulate(ga.ir, env)
            T = x * y + c
            m = (-T * pinv) % 2**256
            assert sum(env["h%d" % k] << (32 * k) for k in range(9)) == (T + m * P) >> 256, "mul_acc mismatch"
    rnd = random.Random(5)
    for x, y in _cases(500, 5, False):
        env = {}
        for k in range(8):
            env["a%d" % k] = (x >> (32 * k)) & M32
            env["b%d" % k] = (y >> (32 * k)) & M32
        simulate(g5.ir, env)
        assert sum(env["t%d" % k] << (32 * k) for k in range(16)) == x * y
    txt = (HEADER + _emit_fn("fp_mul_t9", gm, 2) + _emit_fn("fp_sqr_t9", gs, 1, outs=["h%d" % k for k in range(9)] + ["qx"]) +
           "// cross products (i, j) whose exact high word enters qx above\n" +
           "#define ECB200_SQR_EXACT_PAIRS {%s}\n\n" % ", ".join("{%d, %d}" % ij for ij in gs.exact_pairs) +
           "// t = (a*b + c + m*p) / 2^256 with a 512-bit addend c; h8 in {0, 1, 2}\n" +
           _emit_fn("fp_mul_acc_t9", ga, 2, extra_in=["c%d" % k for k in range(16)]) +
           "// T = a*b, the exact 512-bit product (mul.h:150-158), 16 words\n" +
           _emit_fn("fp_mul512_words", g5, 2, outs=["t%d" % k for k in range(16)]) + "}  // namespace ecb200\n")
    with open(path, "w") as f:
        f.write(txt)
    return nm + ns, gm, gs


def _stats(g):
    wide = sum(1 for i in g.ir if i[0] in ("mulw", "madw"))
    other = sum(1 for i in g.ir if i[0] not in ("mulw", "madw", "endchain"))
    pairs = sum(1 for i in g.ir if i[0] == "addw")
    return wide, other + pairs


if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    if "--check-only" in sys.argv:
        print("mul: simulated cases ok:", check(gen_mul(), False, 20000))
        print("sqr: simulated cases ok:", check(gen_sqr(), True, 20000))
    else:
        n, gm, gs = emit_header(os.path.join(here, "fp256_mul_gen.cuh"))
        print("wrote fp256_mul_gen.cuh; simulated %d cases ok; mul: %d wide multiply-adds + %d ALU ops; sqr: %d + %d"
              % ((n,) + _stats(gm) + _stats(gs)))
