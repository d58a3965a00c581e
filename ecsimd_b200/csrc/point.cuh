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
template <bool QUIRK>
__device__ __forceinline__ jac pt_dblu(jac& P) {
  const fe B = fp_sqr<QUIRK>(P.x);
  const fe E = fp_sqr<QUIRK>(P.y);
  const fe L = fp_sqr<QUIRK>(E);
  const fe S = fp_shl1(fp_sub(fp_sub(fp_sqr<QUIRK>(fp_add(P.x, E)), B), L));
  const fe M = fp_add(fp_add(fp_shl1(B), B), fe_AM());
  jac r;
  r.x = fp_sub(fp_sqr<QUIRK>(M), fp_shl1(S));
  const fe L8 = fp_shl<3>(L);
  r.y = fp_sub(fp_mul(M, fp_sub(S, r.x)), L8);
  r.z = fp_shl1(P.y);
  P.x = S;
  P.y = L8;
  P.z = r.z;
  return r;
}

// ZADDU: returns P + O (same Z required); rewrites P with the Z of the result.
template <bool QUIRK>
__device__ __forceinline__ jac pt_zaddu(jac& P, const jac& O) {
  const fe dx = fp_sub(P.x, O.x);
  const fe dy = fp_sub(P.y, O.y);
  const fe C = fp_sqr<QUIRK>(dx);
  const fe W1 = fp_mul(P.x, C);
  const fe W2 = fp_mul(O.x, C);
  const fe D = fp_sqr<QUIRK>(dy);
  const fe A1 = fp_mul(P.y, fp_sub(W1, W2));
  jac r;
  r.x = fp_sub(fp_sub(D, W1), W2);
  r.y = fp_sub(fp_mul(dy, fp_sub(W1, r.x)), A1);
  r.z = fp_mul(P.z, dx);
  P.x = W1;
  P.y = A1;
  P.z = r.z;
  return r;
}

// ZDAU core on bare coordinates: (X1,Y1) <- 2*(X1,Y1) + (X2,Y2); (X2,Y2) <- the same
// point (X2,Y2) re-scaled to the new common Z; Z <- new Z.
template <bool QUIRK>
__device__ __forceinline__ void pt_zdau_xy(fe& X1, fe& Y1, fe& X2, fe& Y2, fe& Z) {
  const fe dx = fp_sub(X1, X2);
  const fe dy = fp_sub(Y1, Y2);
  const fe Cp = fp_sqr<QUIRK>(dx);
  const fe W1p = fp_mul(X1, Cp);
  const fe W2p = fp_mul(X2, Cp);
  const fe Dp = fp_sqr<QUIRK>(dy);
  const fe A1p = fp_mul(Y1, fp_sub(W1p, W2p));
  const fe X3pc = fp_sub(fp_sub(Dp, W1p), W2p);
  const fe e3 = fp_sub(X3pc, W1p);
  const fe C = fp_sqr<QUIRK>(e3);
  const fe A2 = fp_shl1(A1p);
  // Y3p = ((Y1-Y2) + (W1p-X3pc))^2 - Dp - C - 2*A1p
  const fe Y3p = fp_sub(fp_sub(fp_sub(fp_sqr<QUIRK>(fp_add(dy, fp_sub(W1p, X3pc))), Dp), C), A2);
  const fe W1 = fp_mul(fp_shl<2>(X3pc), C);
  const fe W2 = fp_mul(fp_shl<2>(W1p), C);
  const fe ym = fp_sub(Y3p, A2);
  const fe yp = fp_add(Y3p, A2);
  const fe D = fp_sqr<QUIRK>(ym);
  const fe A1 = fp_mul(Y3p, fp_sub(W1, W2));
  const fe X3 = fp_sub(fp_sub(D, W1), W2);
  const fe Y3 = fp_sub(fp_mul(ym, fp_sub(W1, X3)), A1);
  // Z3 = Z * ((X1 - X2 + X3pc - W1p)^2 - Cp - C)
  const fe Z3 = fp_mul(Z, fp_sub(fp_sub(fp_sqr<QUIRK>(fp_sub(fp_add(dx, X3pc), W1p)), Cp), C));
  const fe Dc = fp_sqr<QUIRK>(yp);
  const fe X2n = fp_sub(fp_sub(Dc, W1), W2);
  const fe Y2n = fp_sub(fp_mul(yp, fp_sub(W1, X2n)), A1);
  X1 = X3; Y1 = Y3;
  X2 = X2n; Y2 = Y2n;
  Z = Z3;
}

// ZDAU(P, Q&): returns 2P + Q, rewrites Q (same point, new Z).  Z of P is used.
template <bool QUIRK>
__device__ __forceinline__ jac pt_zdau(const jac& P, jac& Q) {
  jac r = P;
  pt_zdau_xy<QUIRK>(r.x, r.y, Q.x, Q.y, r.z);
  Q.z = r.z;
  return r;
}

// ADD_Z2_1(A, B): A + B with Z(B) == R assumed (B.z is never read).
template <bool QUIRK>
__device__ __forceinline__ jac pt_add_z2_1(const jac& A, const fe& X2, const fe& Y2) {
  const fe Z1Z1 = fp_sqr<QUIRK>(A.z);
  const fe U2 = fp_mul(X2, Z1Z1);
  const fe S2 = fp_mul(fp_mul(Y2, A.z), Z1Z1);
  const fe H = fp_sub(U2, A.x);
  const fe HH = fp_sqr<QUIRK>(H);
  const fe I = fp_shl<2>(HH);
  const fe J = fp_mul(H, I);
  const fe r = fp_shl1(fp_sub(S2, A.y));
  const fe V = fp_mul(A.x, I);
  jac o;
  o.x = fp_sub(fp_sub(fp_sqr<QUIRK>(r), J), fp_shl1(V));
  o.y = fp_sub(fp_mul(r, fp_sub(V, o.x)), fp_mul(fp_shl1(A.y), J));
  o.z = fp_sub(fp_sub(fp_sqr<QUIRK>(fp_add(A.z, H)), Z1Z1), HH);
  return o;
}

template <bool QUIRK>
__device__ __forceinline__ jac pt_trplu(jac& P) {
  const jac dbl = pt_dblu<QUIRK>(P);
  return pt_zaddu<QUIRK>(P, dbl);
}

// scalar_mult: Joye's right-to-left co-Z double-add ladder with bit 0 forced to 1 and
// a final conditional subtraction of P (curve_group.h:189-218).  k is the raw 256-bit
// scalar (never reduced mod the group order).  The two masked swaps per step of the
// reference (swap.h:47-56) bracket each ZDAU; the closing swap of step b and the
// opening swap of step b+1 are merged into one swap on (bit_b xor bit_{b+1}).
// SYNC: all warps of the block meet at a barrier once per ladder step, so that they walk
// the (large, fully unrolled) loop body together and share instruction-cache lines.
template <bool QUIRK, bool SYNC = false>
__device__ __forceinline__ jac pt_scalar_mult(const uint32_t (&k)[8], const fe& Px, const fe& Py) {
  jac P;
  P.x = Px; P.y = Py; P.z = fe_R();
  const fe oppY = fp_neg(Py);
  jac base = pt_trplu<QUIRK>(P);
  fe Z = base.z;
  // state: (base.x, base.y) and (P.x, P.y) share Z.
  uint32_t prev = (k[0] >> 1) & 1u;  // pending swap
#pragma unroll 1
  for (int b = 2; b < 256; b++) {
    const uint32_t bit = (k[b >> 5] >> (b & 31)) & 1u;
    const uint32_t sw = prev ^ bit;
    fe_cswap(sw, P.x, base.x);
    fe_cswap(sw, P.y, base.y);
    pt_zdau_xy<QUIRK>(base.x, base.y, P.x, P.y, Z);
    prev = bit;
    if (SYNC) __syncthreads();
  }
  fe_cswap(prev, P.x, base.x);
  fe_cswap(prev, P.y, base.y);
  P.z = Z;
  const jac Psub = pt_add_z2_1<QUIRK>(P, Px, oppY);
  const bool odd = (k[0] & 1u) != 0u;
  jac out;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    out.x.v[i] = odd ? P.x.v[i] : Psub.x.v[i];
    out.y.v[i] = odd ? P.y.v[i] : Psub.y.v[i];
    out.z.v[i] = odd ? P.z.v[i] : Psub.z.v[i];
  }
  return out;
}

}  // namespace ecb200
