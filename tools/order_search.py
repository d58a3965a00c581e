#!/usr/bin/env python3
"""Search the statement order of the ZDAU ladder step against ptxas.

The ladder loop is issue-bound: its cost is 4.3 x (#IMAD.WIDE) + ~1.05 x (#other
instructions).  The number of other instructions ptxas emits (register moves, predicate and
register spills, rematerialised multiplies) depends chaotically on the order of the ~45 field
operations of the step, so this tool hill-climbs over valid topological orders, compiling the real
ladder kernel (SOA, per-lane P) for each candidate and counting the SASS of its hot loop.

usage: order_search.py [iterations] [seed]      -> writes ecsimd_b200/csrc/zdau_order.inc when it improves
"""
import os
import random
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCR = os.path.join(ROOT, "build", "scratch")
sys.path.insert(0, os.path.join(ROOT, "tools"))

# (name, defs, uses, text).  fm chains: f1/f2/f3 are threaded through their squarings in order.
STMTS = [
    ("dx", "const fe dx = fp_sub(X1, X2);"),
    ("dy", "const fe dy = fp_sub(Y1, Y2);"),
    ("Cp", "fe Cp = fp_sqr_acc<QUIRK>(dx, md, f1);"),
    ("Dp", "fe Dp = fp_sqr_acc<QUIRK>(dy, md, f1);"),
    ("chk1", "fp_quirk_check<QUIRK>(md, f1, dx, Cp, dy, Dp);"),
    ("W1p", "const fe W1p = fp_mul(X1, Cp, md);"),
    ("W2p", "const fe W2p = fp_mul(X2, Cp, md);"),
    ("dW", "const fe dW = fp_sub(W1p, W2p);"),
    ("A1p", "const fe A1p = fp_mul(Y1, dW, md);"),
    ("t1", "const fe t1 = fp_sub(Dp, W1p);"),
    ("X3pc", "const fe X3pc = fp_sub(t1, W2p);"),
    ("e3", "const fe e3 = fp_sub(X3pc, W1p);"),
    ("de", "const fe de = fp_sub(dy, e3);"),
    ("xe", "const fe xe = fp_add(dx, e3, md);"),
    ("C", "fe C = fp_sqr_acc<QUIRK>(e3, md, f2);"),
    ("s4", "fe s4 = fp_sqr_acc<QUIRK>(de, md, f2);"),
    ("s6", "fe s6 = fp_sqr_acc<QUIRK>(xe, md, f2);"),
    ("chk2", "fp_quirk_check<QUIRK>(md, f2, e3, C, de, s4, xe, s6);"),
    ("C4", "const fe C4 = C4EXPR;"),
    ("W1", "const fe W1 = fp_mul(X3pc, C4, md);"),
    ("W2", "const fe W2 = fp_mul(W1p, C4, md);"),
    ("z1", "const fe z1 = fp_sub(s6, Cp);"),
    ("z2", "const fe z2 = fp_sub(z1, C);"),
    ("Z3", "const fe Z3 = fp_mul(Z, z2, md);"),
    ("A2", "const fe A2 = fp_shl1(A1p, md);"),
    ("y1", "const fe y1 = fp_sub(s4, Dp);"),
    ("yp", "const fe yp = fp_sub(y1, C);"),
    ("Y3p", "const fe Y3p = fp_sub(yp, A2);"),
    ("ym", "const fe ym = fp_sub(Y3p, A2);"),
    ("D", "fe D = fp_sqr_acc<QUIRK>(ym, md, f3);"),
    ("Dc", "fe Dc = fp_sqr_acc<QUIRK>(yp, md, f3);"),
    ("chk3", "fp_quirk_check<QUIRK>(md, f3, ym, D, yp, Dc);"),
    # Y3 = ym*(W1 - X3) - Y3p*(W1 - W2) and Y2n = yp*(W1 - X2n) - Y3p*(W1 - W2) as (a*b + c) with ONE reduction each,
    # sharing the unreduced product TA = Y3p*(W2 - W1)  (fp_mul_wide / fp_mul_acc, fp256.cuh)
    ("dW2n", "const fe dW2n = fp_sub(W2, W1);"),
    ("TA", "const fe512 TA = fp_mul_wide(Y3p, dW2n);"),
    ("W12", "const fe W12 = fp_add(W1, W2, md);"),
    ("X3", "const fe X3 = fp_sub(D, W12);"),
    ("u1", "const fe u1 = fp_sub(W1, X3);"),
    ("Y3", "const fe Y3 = fp_mul_acc(TA, ym, u1, md);"),
    ("X2n", "const fe X2n = fp_sub(Dc, W12);"),
    ("u2", "const fe u2 = fp_sub(W1, X2n);"),
    ("Y2n", "const fe Y2n = fp_mul_acc(TA, yp, u2, md);"),
]
# Equivalent forms of single statements (commutative operands, or the other association of a double
# subtraction): same canonical values, different register pairing for ptxas.  A set of names selects the
# alternative text.
ALT = {
    "W1p": "const fe W1p = fp_mul(Cp, X1, md);",
    "W2p": "const fe W2p = fp_mul(Cp, X2, md);",
    "A1p": "const fe A1p = fp_mul(dW, Y1, md);",
    "W1": "const fe W1 = fp_mul(C4, X3pc, md);",
    "W2": "const fe W2 = fp_mul(C4, W1p, md);",
    "Z3": "const fe Z3 = fp_mul(z2, Z, md);",
    "TA": "const fe512 TA = fp_mul_wide(dW2n, Y3p);",
    "Y3": "const fe Y3 = fp_mul_acc(TA, u1, ym, md);",
    "Y2n": "const fe Y2n = fp_mul_acc(TA, u2, yp, md);",
    "xe": "const fe xe = fp_add(e3, dx, md);",
    "W12": "const fe W12 = fp_add(W2, W1, md);",
    "t1": "const fe t1 = fp_sub(Dp, W2p);",      # with X3pc = t1 - W1p
    "z1": "const fe z1 = fp_sub(s6, C);",        # with z2 = z1 - Cp
    "y1": "const fe y1 = fp_sub(s4, C);",        # with yp = y1 - Dp
}
ALT_PARTNER = {"t1": ("X3pc", "const fe X3pc = fp_sub(t1, W1p);"), "z1": ("z2", "const fe z2 = fp_sub(z1, Cp);"),
               "y1": ("yp", "const fe yp = fp_sub(y1, Dp);")}
NAMES = [n for n, _ in STMTS]
TEXT = dict(STMTS)
INPUTS = {"X1", "Y1", "X2", "Y2", "Z", "md", "QUIRK", "fe", "const", "fp_sub", "fp_add", "fp_mul", "fp_sqr_acc", "fp_quirk_check",
          "fp_shl1", "f1", "f2", "f3", "C4EXPR", "fe512", "fp_mul_wide", "fp_mul_acc"}
DEPS = {}
for n, t in STMTS:
    rhs = t.split("=", 1)[1] if n not in ("chk1", "chk2", "chk3") else t
    ids = set(re.findall(r"[A-Za-z_][A-Za-z0-9_]*", rhs)) - INPUTS - {n}
    DEPS[n] = {i for i in ids if i in TEXT}
# the filter words f1/f2/f3 are read-modify-write: squarings of a group before their check
DEPS["C4"] |= {"C"}
DEPS["t1"] |= {"W2p"}; DEPS["X3pc"] |= {"W1p"}      # union over both associations of Dp - W1p - W2p
DEPS["z1"] |= {"C"}; DEPS["z2"] |= {"Cp"}
DEPS["y1"] |= {"C"}; DEPS["yp"] |= {"Dp"}
DEPS["chk1"] |= {"Cp", "Dp"}
DEPS["chk2"] |= {"C", "s4", "s6"}
DEPS["chk3"] |= {"D", "Dc"}
# the checks repair the squares in place: every consumer of a group's results comes after its check
for grp, chk in ((("Cp", "Dp"), "chk1"), (("C", "s4", "s6"), "chk2"), (("D", "Dc"), "chk3")):
    for n_ in NAMES:
        if n_ != chk and n_ not in grp and DEPS[n_] & set(grp):
            DEPS[n_].add(chk)


def valid(order):
    seen = set()
    for n in order:
        if not DEPS[n] <= seen:
            return False
        seen.add(n)
    return True


def emit(order, c4_lazy, path, alts=frozenset()):
    partner = {ALT_PARTNER[a][0]: ALT_PARTNER[a][1] for a in alts if a in ALT_PARTNER}
    with open(path, "w") as f:
        f.write("// generated by tools/order_search.py: statement order of the ZDAU step chosen against ptxas\n")
        f.write("  QuirkAcc f1, f2, f3;\n")
        for n in order:
            t = ALT[n] if n in alts else partner.get(n, TEXT[n])
            t = t.replace("C4EXPR", "fp_shl2_mulonly(C, md)" if c4_lazy else "fp_shl<2>(C, md)")
            f.write("  " + t + "\n")


TU = r'''
#include "../../ecsimd_b200/csrc/layout.cuh"
#include "../../ecsimd_b200/csrc/point.cuh"
using namespace ecb200;
// the same inputs as k_scalar_mult_sync (kernels_point.cu): the scalar staged in shared memory, the lane index
// recomputed after the loop -- the loop's live set decides how many moves and spills ptxas adds
struct SrcG {
  const void* k; const void* P; size_t n, i; const uint32_t* sk;
  __device__ __forceinline__ uint32_t kword_global(int w) const { return Layout<L_SOA>::load_word(k, n, i, 1, 0, w); }
  __device__ __forceinline__ uint32_t kword(int w) const { return sk[w * blockDim.x]; }
  __device__ __forceinline__ void point(fe& x, fe& y) const {
    const void* p = P; asm volatile("" : "+l"(p));
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; asm volatile("" : "+l"(t));
    const size_t ii = t < n ? t : n - 1;
    x = Layout<L_SOA>::load(p, n, ii, 3, 0); y = Layout<L_SOA>::load(p, n, ii, 3, 1);
  }
  __device__ __forceinline__ void table(uint32_t, fe (&)[5]) const {}
};
extern "C" __global__ void __launch_bounds__(512, 1) k_ladder(void* __restrict__ out, const void* __restrict__ k, const void* __restrict__ P, size_t n) {
  const size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t i = i0 < n ? i0 : n - 1;
  __shared__ uint32_t s_k[8 * 512];
  const SrcG src{k, P, n, i, s_k + threadIdx.x};
#pragma unroll
  for (int w = 0; w < 8; w++) s_k[w * 512 + threadIdx.x] = src.kword_global(w);
  const jac r = pt_scalar_mult<true, true, 0>(src);
  size_t j0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; asm volatile("" : "+l"(j0));
  if (j0 < n) { Layout<L_SOA>::store(out, n, j0, 3, 0, r.x); Layout<L_SOA>::store(out, n, j0, 3, 1, r.y); Layout<L_SOA>::store(out, n, j0, 3, 2, r.z); }
}
'''


def evaluate(order, c4_lazy, tag, alts=frozenset()):
    inc = os.path.join(SCR, "order_%s.inc" % tag)
    emit(order, c4_lazy, inc, alts)
    cu = os.path.join(SCR, "ladder_tu_%s.cu" % tag)
    with open(cu, "w") as f:
        f.write(TU)
    cubin = os.path.join(SCR, "ladder_tu_%s.cubin" % tag)
    r = subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-diag-suppress", "550", "-cubin",
                        "-DECB200_ZDAU_ORDER=2", '-DECB200_ZDAU_ORDER_FILE="%s"' % inc, "-o", cubin, cu], capture_output=True, text=True)
    if r.returncode:
        raise RuntimeError(r.stderr[-2000:])
    sass = subprocess.run(["cuobjdump", "-sass", cubin], capture_output=True, text=True).stdout
    # hot loop = the backward branch range containing BAR.SYNC
    ins = []
    for l in sass.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    bars = [a for a, t in ins if "BAR.SYNC" in t]
    best = None
    for a, t in ins:
        m = re.search(r"BRA.*?(0x[0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a and any(int(m.group(1), 16) <= b <= a for b in bars):
            if best is None or a - int(m.group(1), 16) > best[1] - best[0]:
                best = (int(m.group(1), 16), a)
    loop_ins = [(a, t) for a, t in ins if best[0] <= a <= best[1]]
    # drop the cold blocks (rarely taken slow paths: a forward branch that jumps over a CALL)
    cold = set()
    for a, t in loop_ins:
        m = re.search(r"BRA\s+(?:!?U?P\d+,\s*)?(0x[0-9a-f]+)", t)
        if m and t.startswith("@"):
            tgt = int(m.group(1), 16)
            if tgt > a and any("CALL" in t2 for a2, t2 in loop_ins if a < a2 < tgt):
                cold.update(a2 for a2, t2 in loop_ins if a < a2 < tgt)
    loop = [t for a, t in loop_ins if a not in cold]
    n = len(loop)
    w = sum(1 for t in loop if "IMAD.WIDE" in t)
    mem = sum(1 for t in loop if re.match(r"(@!?P\d+\s+)?(LDL|STL)", t))
    return 4.3 * w + 1.05 * (n - w) + 1.0 * mem, n, w, mem


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    tag = "s%d" % seed
    rnd = random.Random(seed)
    order = list(NAMES)
    c4 = False
    alts = frozenset()
    if len(sys.argv) > 3:   # start from "lazy|canon [alt=a,b,c] name name ..."
        toks = open(sys.argv[3]).read().split()
        c4 = toks[0] == "lazy"
        toks = toks[1:]
        if toks and toks[0].startswith("alt="):
            alts = frozenset(x for x in toks[0][4:].split(",") if x)
            toks = toks[1:]
        order = toks
        assert valid(order) and sorted(order) == sorted(NAMES)
    best = evaluate(order, c4, tag, alts)
    print("start", best, flush=True)
    log = open(os.path.join(SCR, "order_search_%s.log" % tag), "a")
    for it in range(iters):
        cand = list(order)
        cc4, calts = c4, alts
        r = rnd.random()
        if r < 0.05:
            cc4 = not c4
        elif r < 0.5:
            a = rnd.choice(sorted(ALT))
            calts = alts ^ {a}
        else:
            # move one statement to another valid position (a few times)
            for _ in range(rnd.choice([1, 1, 2, 3])):
                for _try in range(50):
                    i = rnd.randrange(len(cand))
                    j = rnd.randrange(len(cand))
                    if i == j:
                        continue
                    c2 = list(cand)
                    x = c2.pop(i)
                    c2.insert(j, x)
                    if valid(c2):
                        cand = c2
                        break
        if cand == order and cc4 == c4 and calts == alts:
            continue
        try:
            sc = evaluate(cand, cc4, tag, calts)
        except RuntimeError as e:
            print("compile error", str(e)[:300])
            continue
        log.write("%d %r %r %s %s\n" % (it, sc, cc4, ",".join(sorted(calts)), " ".join(cand)))
        log.flush()
        if sc[0] <= best[0]:
            if sc[0] < best[0]:
                print("iter %d: %.0f (n=%d mem=%d) c4_lazy=%s alts=%s" % (it, sc[0], sc[1], sc[3], cc4, ",".join(sorted(calts))), flush=True)
            order, c4, alts, best = cand, cc4, calts, sc
            emit(order, c4, os.path.join(SCR, "best_order_%s.inc" % tag), alts)
            with open(os.path.join(SCR, "best_order_%s.txt" % tag), "w") as f:
                f.write(("lazy " if c4 else "canon ") + "alt=" + ",".join(sorted(alts)) + " " + " ".join(order) + "\n")
    print("best", best, c4, sorted(alts), " ".join(order))


if __name__ == "__main__":
    main()
