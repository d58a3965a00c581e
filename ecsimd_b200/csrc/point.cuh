// point.cuh -- co-Z Jacobian point arithmetic on P-256 (Montgomery-form coordinates),
// one point per thread, everything in registers.
//
// Formulas: Goundar-Joye-Miyaji co-Z (eprint 2010/309) exactly as instantiated by the
// reference in include/ecsimd/curve_group.h:
//   pt_dblu     = curve_group::DBLU      :64-87
//   pt_zaddu    = curve_group::ZADDU     :91-116
//   pt_zdau     = curve_group::ZDAU      :120-153
//   pt_add_z2_1 = curve_group::ADD_Z2_1  :155-179
//   pt_trplu    = curve_group::TRPLU     :183-186
//   pt_scalar_mult = curve_group::scalar_mult :189-218 (= lib/scalar_mult_p256.cpp:12-14)
// The set of values that get SQUARED is kept identical to the reference (the
// squaring defect reproduced by fp_sqr depends on the operand); everything else is
// free to share sub-expressions, because every field result is a pure function of
// canonical inputs.
#pragma once
#include "fp256.cuh"

namespace ecb200 {

struct jac {
  fe x, y, z;
};

// DBLU: returns 2P; rewrites P as the same point with Z(P) = Z(2P).  Input Z is ignored
// (the reference requires Z == R and never reads it).
template <bool QUIRK, class MD>
__device__ __forceinline__ jac pt_dblu(jac& P, MD& md) {
  const fe B = fp_sqr<QUIRK>(P.x, md);
  const fe E = fp_sqr<QUIRK>(P.y, md);
  const fe L = fp_sqr<QUIRK>(E, md);
  const fe S = fp_shl1(fp_sub(fp_sub(fp_sqr<QUIRK>(fp_add(P.x, E, md), md), B), L), md);
  const fe M = fp_add(fp_add(fp_shl1(B, md), B, md), fe_AM(), md);
  jac r;
  r.x = fp_sub(fp_sqr<QUIRK>(M, md), fp_shl1(S, md));
  const fe L8 = fp_shl<3>(L, md);
  r.y = fp_sub(fp_mul(M, fp_sub(S, r.x), md), L8);
  r.z = fp_shl1(P.y, md);
  P.x = S;
  P.y = L8;
  P.z = r.z;
  return r;
}

// ZADDU: returns P + O (same Z required); rewrites P with the Z of the result.
template <bool QUIRK, class MD>
__device__ __forceinline__ jac pt_zaddu(jac& P, const jac& O, MD& md) {
  const fe dx = fp_sub(P.x, O.x);
  const fe dy = fp_sub(P.y, O.y);
  const fe C = fp_sqr<QUIRK>(dx, md);
  const fe W1 = fp_mul(P.x, C, md);
  const fe W2 = fp_mul(O.x, C, md);
  const fe D = fp_sqr<QUIRK>(dy, md);
  const fe A1 = fp_mul(P.y, fp_sub(W1, W2), md);
  jac r;
  r.x = fp_sub(fp_sub(D, W1), W2);
  r.y = fp_sub(fp_mul(dy, fp_sub(W1, r.x), md), A1);
  r.z = fp_mul(P.z, dx, md);
  P.x = W1;
  P.y = A1;
  P.z = r.z;
  return r;
}

// ZDAU core on bare coordinates: (X1,Y1) <- 2*(X1,Y1) + (X2,Y2); (X2,Y2) <- the same
// point (X2,Y2) re-scaled to the new common Z; Z <- new Z.
#ifndef ECB200_ZDAU_ORDER
#define ECB200_ZDAU_ORDER 2
#endif
#ifndef ECB200_ZDAU_ORDER_FILE
#define ECB200_ZDAU_ORDER_FILE "zdau_order.inc"
#endif
template <bool QUIRK, class MD, int MIDSYNC = 0>
__device__ __forceinline__ void pt_zdau_xy(fe& X1, fe& Y1, fe& X2, fe& Y2, fe& Z, MD& md, int grp = 0) {
#if ECB200_ZDAU_ORDER == 0
  const fe dx = fp_sub(X1, X2);
  const fe dy = fp_sub(Y1, Y2);
  const fe Cp = fp_sqr<QUIRK>(dx, md);
  const fe W1p = fp_mul(X1, Cp, md);
  const fe W2p = fp_mul(X2, Cp, md);
  const fe Dp = fp_sqr<QUIRK>(dy, md);
  const fe A1p = fp_mul(Y1, fp_sub(W1p, W2p), md);
  const fe X3pc = fp_sub(fp_sub(Dp, W1p), W2p);
  const fe e3 = fp_sub(X3pc, W1p);
  const fe C = fp_sqr<QUIRK>(e3, md);
  const fe A2 = fp_shl1(A1p, md);
  // Y3p = ((Y1-Y2) + (W1p-X3pc))^2 - Dp - C - 2*A1p        [(W1p - X3pc) = -e3]
  const fe Y3p = fp_sub(fp_sub(fp_sub(fp_sqr<QUIRK>(fp_sub(dy, e3), md), Dp), C), A2);
  // W1 = 4*X3pc*C, W2 = 4*W1p*C: quadruple C once instead of each multiplicand
  const fe C4 = fp_shl2_mulonly(C, md);
  const fe W1 = fp_mul(X3pc, C4, md);
  const fe W2 = fp_mul(W1p, C4, md);
  const fe W12 = fp_add(W1, W2, md);
  if (MIDSYNC) { if (grp == 1) asm volatile("bar.sync 0;" ::: "memory"); }
  const fe ym = fp_sub(Y3p, A2);
  const fe yp = fp_add(Y3p, A2, md);
  const fe D = fp_sqr<QUIRK>(ym, md);
  const fe A1 = fp_mul(Y3p, fp_sub(W1, W2), md);
  const fe X3 = fp_sub(D, W12);
  const fe Y3 = fp_sub(fp_mul(ym, fp_sub(W1, X3), md), A1);
  // Z3 = Z * ((X1 - X2 + X3pc - W1p)^2 - Cp - C)           [X3pc - W1p = e3]
  const fe Z3 = fp_mul(Z, fp_sub(fp_sub(fp_sqr<QUIRK>(fp_add(dx, e3, md), md), Cp), C), md);
  const fe Dc = fp_sqr<QUIRK>(yp, md);
  const fe X2n = fp_sub(Dc, W12);
  const fe Y2n = fp_sub(fp_mul(yp, fp_sub(W1, X2n), md), A1);
#elif ECB200_ZDAU_ORDER == 2
#include ECB200_ZDAU_ORDER_FILE
#else
  // Same values, statement order chosen so that consecutive multiplications are independent
  // (ptxas then overlaps their carry chains), and the squarings grouped so that one branch
  // resolves the squaring-defect filter of a whole group (fp_sqr_acc / fp_quirk_check).
  const fe dx = fp_sub(X1, X2);
  const fe dy = fp_sub(Y1, Y2);
  QuirkAcc f1;
  fe Cp = fp_sqr_acc<QUIRK>(dx, md, f1);
  fe Dp = fp_sqr_acc<QUIRK>(dy, md, f1);
  fp_quirk_check<QUIRK>(md, f1, dx, Cp, dy, Dp);
  const fe W1p = fp_mul(X1, Cp, md);
  const fe W2p = fp_mul(X2, Cp, md);
  const fe A1p = fp_mul(Y1, fp_sub(W1p, W2p), md);
  const fe X3pc = fp_sub(fp_sub(Dp, W1p), W2p);
  const fe e3 = fp_sub(X3pc, W1p);
  const fe de = fp_sub(dy, e3);                            // (Y1-Y2) + (W1p-X3pc), (W1p - X3pc) = -e3
  const fe xe = fp_add(dx, e3, md);                        // X1 - X2 + X3pc - W1p
  QuirkAcc f2;
  fe C = fp_sqr_acc<QUIRK>(e3, md, f2);
  fe s4 = fp_sqr_acc<QUIRK>(de, md, f2);
  fe s6 = fp_sqr_acc<QUIRK>(xe, md, f2);
  fp_quirk_check<QUIRK>(md, f2, e3, C, de, s4, xe, s6);
  const fe C4 = fp_shl2_mulonly(C, md);                          // W1 = 4*X3pc*C, W2 = 4*W1p*C: quadruple C once
  const fe W1 = fp_mul(X3pc, C4, md);
  const fe W2 = fp_mul(W1p, C4, md);
  const fe Z3 = fp_mul(Z, fp_sub(fp_sub(s6, Cp), C), md);
  if (MIDSYNC) { if (grp == 1) asm volatile("bar.sync 0;" ::: "memory"); }
  const fe A2 = fp_shl1(A1p, md);
  // Y3p = s4 - Dp - C - 2*A1p, yp = Y3p + A2, ym = Y3p - A2: yp is the partial sum itself
  const fe yp = fp_sub(fp_sub(s4, Dp), C);
  const fe Y3p = fp_sub(yp, A2);
  const fe ym = fp_sub(Y3p, A2);
  QuirkAcc f3;
  fe D = fp_sqr_acc<QUIRK>(ym, md, f3);
  fe Dc = fp_sqr_acc<QUIRK>(yp, md, f3);
  fp_quirk_check<QUIRK>(md, f3, ym, D, yp, Dc);
  const fe A1 = fp_mul(Y3p, fp_sub(W1, W2), md);
  const fe W12 = fp_add(W1, W2, md);
  const fe X3 = fp_sub(D, W12);
  const fe Y3 = fp_sub(fp_mul(ym, fp_sub(W1, X3), md), A1);
  const fe X2n = fp_sub(Dc, W12);
  const fe Y2n = fp_sub(fp_mul(yp, fp_sub(W1, X2n), md), A1);
#endif
  X1 = X3; Y1 = Y3;
  X2 = X2n; Y2 = Y2n;
  Z = Z3;
}

// ZDAU(P, Q&): returns 2P + Q, rewrites Q (same point, new Z).  Z of P is used.
template <bool QUIRK, class MD>
__device__ __forceinline__ jac pt_zdau(const jac& P, jac& Q, MD& md) {
  jac r = P;
  pt_zdau_xy<QUIRK>(r.x, r.y, Q.x, Q.y, r.z, md);
  Q.z = r.z;
  return r;
}

// ADD_Z2_1(A, B): A + B with Z(B) == R assumed (B.z is never read).
template <bool QUIRK, class MD>
__device__ __forceinline__ jac pt_add_z2_1(const jac& A, const fe& X2, const fe& Y2, MD& md) {
  const fe Z1Z1 = fp_sqr<QUIRK>(A.z, md);
  const fe U2 = fp_mul(X2, Z1Z1, md);
  const fe S2 = fp_mul(fp_mul(Y2, A.z, md), Z1Z1, md);
  const fe H = fp_sub(U2, A.x);
  const fe HH = fp_sqr<QUIRK>(H, md);
  const fe I = fp_shl<2>(HH, md);
  const fe J = fp_mul(H, I, md);
  const fe r = fp_shl1(fp_sub(S2, A.y), md);
  const fe V = fp_mul(A.x, I, md);
  jac o;
  o.x = fp_sub(fp_sub(fp_sqr<QUIRK>(r, md), J), fp_shl1(V, md));
  o.y = fp_sub(fp_mul(r, fp_sub(V, o.x), md), fp_mul(fp_shl1(A.y, md), J, md));
  o.z = fp_sub(fp_sub(fp_sqr<QUIRK>(fp_add(A.z, H, md), md), Z1Z1), HH);
  return o;
}

template <bool QUIRK, class MD>
__device__ __forceinline__ jac pt_trplu(jac& P, MD& md) {
  const jac dbl = pt_dblu<QUIRK>(P, md);
  return pt_zaddu<QUIRK>(P, dbl, md);
}

// scalar_mult: Joye's right-to-left co-Z double-add ladder with bit 0 forced to 1 and
// a final conditional subtraction of P (curve_group.h:189-218).  k is the raw 256-bit
// scalar (never reduced mod the group order).  The two masked swaps per step of the
// reference (swap.h:47-56) bracket each ZDAU; the closing swap of step b and the
// opening swap of step b+1 are merged into one swap on (bit_b xor bit_{b+1}).
// SYNC: all warps of the block meet at a barrier once per ladder step, so that they walk
// the (large, fully unrolled) loop body together and share instruction-cache lines.
//
// Inputs come through a source object instead of registers: the loop needs one scalar word every
// 32 steps and P only before and after it, so holding them (24 registers) through 254 steps only
// costs spills.  Src provides
//     uint32_t kword(int w) const;                 word w of the scalar
//     void point(fe& x, fe& y) const;              P (affine Montgomery coordinates, Z = R)
//     void table(uint32_t idx, fe (&st)[5]) const; ladder state after bits 1..TABW (TABW > 0 only)
// TABW > 0 (fixed base point): the state (base.x, base.y, P.x, P.y, Z) after the steps for bits
// 1..TABW depends only on those bits, so it is looked up and the ladder resumes at bit TABW+1:
// same values as the full ladder, because the ladder is right-to-left.
struct SrcRegs {
  const uint32_t* k;
  const uint32_t* xy;  // x words 0..7, y words 8..15
  __device__ __forceinline__ uint32_t kword(int w) const { return k[w]; }
  __device__ __forceinline__ void point(fe& x, fe& y) const {
#pragma unroll
    for (int i = 0; i < 8; i++) { x.v[i] = xy[i]; y.v[i] = xy[8 + i]; }
  }
  __device__ __forceinline__ void table(uint32_t, fe (&)[5]) const {}
};

template <bool QUIRK, bool SYNC, int TABW, class MD, class Src>
__device__ __forceinline__ jac pt_scalar_mult_mode(const Src& src, MD& md) {
  static_assert(TABW >= 0 && TABW <= 30, "table index comes from the low scalar word");
  fe bx, by, qx, qy, Z;  // (bx, by) = base, (qx, qy) = P of the reference's loop; they share Z
  const uint32_t k0 = src.kword(0);
  uint32_t prev, w;
  int b0;
  if (TABW == 0) {
    jac P;
    src.point(P.x, P.y);
    P.z = fe_R();
    const jac base = pt_trplu<QUIRK>(P, md);
    bx = base.x; by = base.y; qx = P.x; qy = P.y; Z = base.z;
    prev = (k0 >> 1) & 1u;  // pending swap
    w = k0 >> 2;
    b0 = 2;
  } else {
    fe st[5];
    src.table((k0 >> 1) & ((1u << TABW) - 1u), st);
    bx = st[0]; by = st[1]; qx = st[2]; qy = st[3]; Z = st[4];
    prev = (k0 >> TABW) & 1u;
    b0 = TABW + 1;
    w = src.kword(b0 >> 5) >> (b0 & 31);
  }
#pragma unroll 1
  for (int b = b0; b < 256; b++) {
    if ((b & 31) == 0) w = src.kword(b >> 5);
    const uint32_t bit = w & 1u;
    w >>= 1;
    const uint32_t sw = prev ^ bit;
    fe_cswap(sw, qx, bx);
    fe_cswap(sw, qy, by);
    pt_zdau_xy<QUIRK>(bx, by, qx, qy, Z, md);
    prev = bit;
#ifndef ECB200_SYNC_EVERY
#define ECB200_SYNC_EVERY 1
#endif
    if (SYNC && (ECB200_SYNC_EVERY == 1 || (b % ECB200_SYNC_EVERY) == 0)) __syncthreads();
  }
  fe_cswap(prev, qx, bx);
  fe_cswap(prev, qy, by);
  jac P;
  P.x = qx; P.y = qy; P.z = Z;
  fe Px, Py;
  src.point(Px, Py);
  const jac Psub = pt_add_z2_1<QUIRK>(P, Px, fp_neg(Py), md);
  const bool odd = (src.kword(0) & 1u) != 0u;
  jac out;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    out.x.v[i] = odd ? P.x.v[i] : Psub.x.v[i];
    out.y.v[i] = odd ? P.y.v[i] : Psub.y.v[i];
    out.z.v[i] = odd ? P.z.v[i] : Psub.z.v[i];
  }
  return out;
}

// The exact re-run for lanes the fast ladder flagged (only the 2^-32 cases of the conditional
// subtractions, now that defect squares are repaired in place): same ladder, every rare case
// resolved in place.  Kept out of line so that it costs the hot path nothing.
template <bool QUIRK>
__device__ __noinline__ void pt_scalar_mult_exact(uint32_t* out24, const uint32_t* k8, const uint32_t* xy16) {
  Exact md;
  const SrcRegs src{k8, xy16};
  const jac r = pt_scalar_mult_mode<QUIRK, false, 0>(src, md);
  for (int i = 0; i < 8; i++) { out24[i] = r.x.v[i]; out24[8 + i] = r.y.v[i]; out24[16 + i] = r.z.v[i]; }
}

// Ladder state after the steps for bits 1..tabw of the scalar `bits << 1` on point (px, py), every
// rare case resolved in place: one entry of the fixed-base table.
template <bool QUIRK>
__device__ __noinline__ void pt_ladder_prefix_exact(uint32_t* st40, uint32_t bits, int tabw, const uint32_t* xy16) {
  Exact md;
  jac P;
  for (int i = 0; i < 8; i++) { P.x.v[i] = xy16[i]; P.y.v[i] = xy16[8 + i]; }
  P.z = fe_R();
  const jac base = pt_trplu<QUIRK>(P, md);
  fe bx = base.x, by = base.y, qx = P.x, qy = P.y, Z = base.z;
  uint32_t prev = bits & 1u;  // bit 1 of the scalar
  uint32_t w = bits >> 1;
#pragma unroll 1
  for (int b = 2; b <= tabw; b++) {
    const uint32_t bit = w & 1u;
    w >>= 1;
    const uint32_t sw = prev ^ bit;
    fe_cswap(sw, qx, bx);
    fe_cswap(sw, qy, by);
    pt_zdau_xy<QUIRK>(bx, by, qx, qy, Z, md);
    prev = bit;
  }
  for (int i = 0; i < 8; i++) {
    st40[i] = bx.v[i]; st40[8 + i] = by.v[i]; st40[16 + i] = qx.v[i]; st40[24 + i] = qy.v[i]; st40[32 + i] = Z.v[i];
  }
}

// Fast ladder (Lazy mode) with the exact re-run of flagged lanes.
template <bool QUIRK, bool SYNC, int TABW, class Src>
__device__ __forceinline__ jac pt_scalar_mult(const Src& src) {
  Lazy md;
  jac r = pt_scalar_mult_mode<QUIRK, SYNC, TABW>(src, md);
  if (__builtin_expect(md.flagged(), 0)) {
    uint32_t kk[8], xy[16], o[24];
    fe px, py;
    src.point(px, py);
    for (int i = 0; i < 8; i++) { kk[i] = src.kword(i); xy[i] = px.v[i]; xy[8 + i] = py.v[i]; }
    pt_scalar_mult_exact<QUIRK>(o, kk, xy);
    for (int i = 0; i < 8; i++) { r.x.v[i] = o[i]; r.y.v[i] = o[8 + i]; r.z.v[i] = o[16 + i]; }
  }
  return r;
}

}  // namespace ecb200
