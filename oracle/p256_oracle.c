/* oracle/p256_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C (scalar, one lane at a time) restatement of the aguinet/ecsimd
 * P-256 hot path.  Every function cites the reference file:line it follows
 * (paths relative to the reference checkout).  It exists so that the CUDA
 * engine can be checked on a GPU box where the reference itself is absent.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_ref.py checks every function
 * here bit-for-bit against the reference's own code compiled from
 * /root/reference (oracle/ref_harness.cpp -> oracle/_ref/libecsimd_ref.so)
 * on random, edge and squaring-quirk inputs; tests/test_oracle_golden.py
 * checks it against the committed fixtures in tests/golden/ (generated from
 * the reference by tests/golden/make_golden.py) and against the known-answer
 * vectors of the reference's tests/curve_group.cpp and tests/curve_point.cpp.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product path
 * (ecsimd_b200/csrc) never calls it.
 */
#include "p256_oracle.h"

#include <pthread.h>
#include <string.h>

typedef uint32_t u32;
typedef uint64_t u64;

#define M32 0xffffffffull

/* p = 2^256 - 2^224 + 2^192 + 2^96 - 1, curve_nist_p256.h:16-18; LS word first */
static const u32 P256[8] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0x00000000u,
                            0x00000000u, 0x00000000u, 0x00000001u, 0xffffffffu};
/* R mod p, R^2 mod p, (p-1)R mod p: mgry_csts.h:20-24 (values dumped from the
 * reference; re-checked by tests via ref_constants) */
static const u32 R_P[8] = {0x00000001u, 0x00000000u, 0x00000000u, 0xffffffffu,
                           0xffffffffu, 0xffffffffu, 0xfffffffeu, 0x00000000u};
static const u32 RSQ_P[8] = {0x00000003u, 0x00000000u, 0xffffffffu, 0xfffffffbu,
                             0xfffffffeu, 0xffffffffu, 0xfffffffdu, 0x00000004u};
static const u32 PM1_R_P[8] = {0xfffffffeu, 0xffffffffu, 0xffffffffu, 0x00000001u,
                               0x00000000u, 0x00000000u, 0x00000002u, 0xfffffffeu};
/* Am = to_mgry(-3), Bm = to_mgry(b): curve_group.h:31-32 */
static const u32 AM[8] = {0xfffffffcu, 0xffffffffu, 0xffffffffu, 0x00000003u,
                          0x00000000u, 0x00000000u, 0x00000004u, 0xfffffffcu};
static const u32 BM[8] = {0x29c4bddfu, 0xd89cdf62u, 0x78843090u, 0xacf005cdu,
                          0xf7212ed6u, 0xe5a220abu, 0x04874834u, 0xdc30061du};
/* G in classical affine form: curve_nist_p256.h:26-31 */
static const u32 GX[8] = {0xd898c296u, 0xf4a13945u, 0x2deb33a0u, 0x77037d81u,
                          0x63a440f2u, 0xf8bce6e5u, 0xe12c4247u, 0x6b17d1f2u};
static const u32 GY[8] = {0x37bf51f5u, 0xcbb64068u, 0x6b315eceu, 0x2bce3357u,
                          0x7c0f9e16u, 0x8ee7eb4au, 0xfe1a7f9bu, 0x4fe342e2u};
/* m' = -p^-1 mod 2^32 = 1: mgry_mul.h:33-40 */
#define MPRIME 1u

static __thread u64 g_cnt[6];

int orc_abi_version(void) { return 1; }
void orc_counters_reset(void) { memset(g_cnt, 0, sizeof g_cnt); }
void orc_counters_get(u64 out[6]) { memcpy(out, g_cnt, sizeof g_cnt); }

/* ---- multi-word helpers ------------------------------------------------- */
/* add.h:11-35  r = a + b mod 2^256, returns carry-out */
static u32 add8(u32 r[8], const u32 a[8], const u32 b[8]) {
  u64 c = 0;
  for (int i = 0; i < 8; i++) { c += (u64)a[i] + b[i]; r[i] = (u32)c; c >>= 32; }
  return (u32)c;
}
/* sub.h:12-38  r = a - b mod 2^256, returns borrow-out */
static u32 sub8(u32 r[8], const u32 a[8], const u32 b[8]) {
  u64 bw = 0;
  for (int i = 0; i < 8; i++) { u64 d = (u64)a[i] - b[i] - bw; r[i] = (u32)d; bw = (d >> 32) & 1; }
  return (u32)bw;
}

/* modular.h:10-15 with sub_if_above (sub.h:46-69): the sum is kept iff
 * (sum - p borrows) AND (the addition did not carry); otherwise sum - p. */
void orc1_mod_add(u32 r[8], const u32 a[8], const u32 b[8]) {
  u32 s[8], d[8];
  u32 c = add8(s, a, b);
  u32 bw = sub8(d, s, P256);
  memcpy(r, (bw && !c) ? s : d, 32);
  g_cnt[2]++;
}
/* modular.h:24-41: diff = a-b; if it borrowed, diff + p (mod 2^256). */
void orc1_mod_sub(u32 r[8], const u32 a[8], const u32 b[8]) {
  u32 d[8], da[8];
  u32 bw = sub8(d, a, b);
  add8(da, d, P256);
  memcpy(r, bw ? da : d, 32);
  g_cnt[3]++;
}
/* modular.h:17-22 with shift.h:13-32: 1-bit left shift, carry = old top bit */
void orc1_mod_shl1(u32 r[8], const u32 a[8]) {
  u32 s[8], d[8];
  u32 c = a[7] >> 31;
  for (int i = 7; i > 0; i--) s[i] = (a[i] << 1) | (a[i - 1] >> 31);
  s[0] = a[0] << 1;
  u32 bw = sub8(d, s, P256);
  memcpy(r, (bw && !c) ? s : d, 32);
  g_cnt[4]++;
}

/* mul.h:115-148 (mul_u32_zext) + :150-158: exact 256x256->512 schoolbook on
 * 32-bit digits; the running sums never overflow 64 bits. */
void orc1_mul512(u32 r[16], const u32 a[8], const u32 b[8]) {
  u64 ret[16];
  memset(ret, 0, sizeof ret);
  for (int i = 0; i < 8; i++) {
    u64 highprev = 0;
    for (int j = 0; j < 8; j++) {
      u64 t = (u64)a[i] * b[j];
      t += ret[i + j];
      t += highprev;
      ret[i + j] = t & M32;
      highprev = t >> 32;
    }
    ret[i + 8] = highprev;
  }
  for (int k = 0; k < 16; k++) r[k] = (u32)ret[k];
}

/* mul.h:160-212 (square_u32_zext) + :214-221, INCLUDING its defect: the
 * doubled cross product `t <<= 1; t += ret; t += prevs[0]` (mul.h:192-195) is
 * evaluated in wrap-around 64-bit lanes, so when (2*pr mod 2^64) + ret + prev
 * reaches 2^64 the carry is silently dropped.  `wraps` counts those events. */
static int square512_impl(u32 r[16], const u32 a[8]) {
  u64 ret[16];
  int wraps = 0;
  memset(ret, 0, sizeof ret);
  for (int i = 0; i < 8; i++) {
    u64 t = (u64)a[i] * a[i];
    t += ret[2 * i];
    ret[2 * i] = t & M32;
    u64 prevs0 = t >> 32, prevs1 = 0;
    for (int j = i + 1; j < 8; j++) {
      u64 pr = (u64)a[i] * a[j];
      u64 carry = pr >> 63;
      u64 t2 = pr << 1;
      u64 s1 = t2 + ret[i + j];
      if (s1 < t2) wraps++;
      u64 s2 = s1 + prevs0;
      if (s2 < s1) wraps++;
      ret[i + j] = s2 & M32;
      prevs0 = prevs1;
      prevs0 += s2 >> 32;
      prevs1 = carry;
    }
    ret[i + 8] += prevs0; /* mul.h:206 "TODO: carry?" */
    if (i + 9 < 16) ret[i + 9] = prevs1;
  }
  /* trunc_u64x32, mul.h:85-113: keeps the low 32 bits of every digit */
  for (int k = 0; k < 16; k++) r[k] = (u32)ret[k];
  return wraps;
}
void orc1_square512(u32 r[16], const u32 a[8]) { g_cnt[5] += (u64)square512_impl(r, a); }
int orc1_square_quirk_hits(const u32 a[8]) { u32 r[16]; return square512_impl(r, a); }

/* mgry_mul.h:84-121 (mgry_reduce): word-serial Montgomery reduction on 32-bit
 * digits held in 64-bit lanes.  limb_mul_zext = mul.h:223-252,
 * add_no_carry_u32_zext = mgry_mul.h:52-82, final step sub_if_above<4> on the
 * 9-digit quotient (sub.h:46-69). */
void orc1_mgry_reduce(u32 r[8], const u32 t[16]) {
  u64 acc[17];
  for (int k = 0; k < 16; k++) acc[k] = t[k];
  acc[16] = 0;
  for (int i = 0; i < 8; i++) {
    u32 m = (u32)((acc[i] & M32) * MPRIME); /* mullow keeps 32x32 (mgry_mul.h:113) */
    u64 prod[9], highprev = 0;
    for (int k = 0; k < 8; k++) {
      u64 x = (u64)P256[k] * m + highprev;
      prod[k] = x & M32;
      highprev = x >> 32;
    }
    prod[8] = highprev & M32;
    for (int k = 0; k < 9 && i + k < 17; k++) acc[i + k] += prod[k];
    for (int k = 1; k < 17; k++) { acc[k] += acc[k - 1] >> 32; acc[k - 1] &= M32; }
    acc[16] &= M32;
  }
  /* quotient = digits 8..16 (9 digits, fits 257 bits); subtract p iff >= p */
  u64 bw = 0;
  u32 d[9];
  for (int k = 0; k < 9; k++) {
    u64 pk = (k < 8) ? P256[k] : 0;
    u64 x = (acc[8 + k] & M32) - pk - bw;
    d[k] = (u32)x;
    bw = (x >> 32) & 1;
  }
  for (int k = 0; k < 8; k++) r[k] = bw ? (u32)acc[8 + k] : d[k];
}

/* mgry_ops.h:31-35 */
static void fmul(u32 r[8], const u32 a[8], const u32 b[8]) {
  u32 t[16];
  orc1_mul512(t, a, b);
  orc1_mgry_reduce(r, t);
  g_cnt[0]++;
}
/* mgry_ops.h:37-42 */
static void fsqr(u32 r[8], const u32 a[8]) {
  u32 t[16];
  orc1_square512(t, a);
  orc1_mgry_reduce(r, t);
  g_cnt[1]++;
}
#define fadd orc1_mod_add /* mgry_ops.h:10-13 */
#define fsub orc1_mod_sub /* mgry_ops.h:26-29 */
/* mgry_ops.h:15-24 */
static void fshl(u32 r[8], const u32 a[8], int count) {
  u32 t[8];
  memcpy(t, a, 32);
  for (int i = 0; i < count; i++) { u32 u[8]; orc1_mod_shl1(u, t); memcpy(t, u, 32); }
  memcpy(r, t, 32);
}
/* gfp.h:60-64 */
static void fopp(u32 r[8], const u32 a[8]) {
  u32 t[8];
  fsub(t, a, R_P);
  fsub(r, PM1_R_P, t);
}
/* mgry.h:47-50 */
static void from_classical(u32 r[8], const u32 a[8]) {
  u32 t[16];
  orc1_mul512(t, a, RSQ_P);
  orc1_mgry_reduce(r, t);
}
/* mgry.h:52-55 */
static void to_classical(u32 r[8], const u32 a[8]) {
  u32 t[16];
  memset(t, 0, sizeof t);
  memcpy(t, a, 32);
  orc1_mgry_reduce(r, t);
}
/* mgry_ops.h:44-86: LSB-first square-and-multiply.  The reference walks the
 * u64 limbs below the top non-zero one bit by bit (always squaring), then the
 * top limb until it is exhausted (no square after the last bit). */
static void fpow(u32 r[8], const u32 a[8], const u32 e[8]) {
  int top = -1;
  for (int b = 255; b >= 0; b--) if ((e[b >> 5] >> (b & 31)) & 1) { top = b; break; }
  u32 res[8], base[8], t[8];
  memcpy(res, R_P, 32);
  memcpy(base, a, 32);
  for (int b = 0; b <= top; b++) {
    if ((e[b >> 5] >> (b & 31)) & 1) { fmul(t, res, base); memcpy(res, t, 32); }
    if (b < top) { fsqr(t, base); memcpy(base, t, 32); }
  }
  memcpy(r, res, 32);
}
/* gfp.h:42-44, exponent p-2 (gfp.h:80-81) */
static void finv(u32 r[8], const u32 a[8]) {
  u32 e[8];
  memcpy(e, P256, 32);
  e[0] -= 2;
  fpow(r, a, e);
}
/* gfp.h:46-54, exponent (p+1)/4 (gfp.h:85-87); returns 1 iff r^2 == a */
static int fsqrt(u32 r[8], const u32 a[8]) {
  /* (p+1)/4 = 2^254 - 2^222 + 2^190 + 2^94 */
  static const u32 e[8] = {0, 0, 0x40000000u, 0, 0, 0x40000000u, 0xc0000000u, 0x3fffffffu};
  u32 chk[8];
  fpow(r, a, e);
  fsqr(chk, r);
  return memcmp(chk, a, 32) == 0;
}

/* ---- points: X|Y|Z, 24 words ------------------------------------------------ */
typedef struct { u32 x[8], y[8], z[8]; } jac;

/* curve_group.h:64-87 */
static void DBLU(jac* ret, jac* P) {
  u32 B[8], E[8], L[8], S[8], Mv[8], t[8], u[8], Lm8[8];
  fsqr(B, P->x);
  fsqr(E, P->y);
  fsqr(L, E);
  fadd(t, P->x, E); fsqr(u, t); fsub(t, u, B); fsub(u, t, L); fshl(S, u, 1);
  fshl(t, B, 1); fadd(u, t, B); fadd(Mv, u, AM);
  fsqr(t, Mv); fshl(u, S, 1); fsub(ret->x, t, u);
  fshl(Lm8, L, 3);
  fsub(t, S, ret->x); fmul(u, Mv, t); fsub(ret->y, u, Lm8);
  fshl(ret->z, P->y, 1);
  memcpy(P->x, S, 32);
  memcpy(P->y, Lm8, 32);
  memcpy(P->z, ret->z, 32);
}
/* curve_group.h:91-116 */
static void ZADDU(jac* ret, jac* P, const jac* O) {
  u32 C[8], W1[8], W2[8], D[8], A1[8], t[8], u[8], dx[8], dy[8];
  fsub(dx, P->x, O->x); fsqr(C, dx);
  fmul(W1, P->x, C);
  fmul(W2, O->x, C);
  fsub(dy, P->y, O->y); fsqr(D, dy);
  fsub(t, W1, W2); fmul(A1, P->y, t);
  fsub(t, D, W1); fsub(ret->x, t, W2);
  fsub(dy, P->y, O->y); fsub(t, W1, ret->x); fmul(u, dy, t); fsub(ret->y, u, A1);
  fsub(dx, P->x, O->x); fmul(ret->z, P->z, dx);
  memcpy(P->x, W1, 32);
  memcpy(P->y, A1, 32);
  memcpy(P->z, ret->z, 32);
}
/* curve_group.h:120-153; sub-expressions are recomputed exactly where the
 * source recomputes them so the op counters match the reference. */
static void ZDAU(jac* ret, const jac* P, jac* Q) {
  const u32 *X1 = P->x, *Y1 = P->y, *Z = P->z;
  u32 *X2 = Q->x, *Y2 = Q->y;
  u32 Cp[8], W1p[8], W2p[8], Dp[8], A1p[8], X3pc[8], C[8], Y3p[8], W1[8], W2[8], D[8], A1[8], Dc[8];
  u32 t[8], u[8], v[8], w[8];
  fsub(t, X1, X2); fsqr(Cp, t);                                  /* :129 */
  fmul(W1p, X1, Cp);                                             /* :130 */
  fmul(W2p, X2, Cp);                                             /* :131 */
  fsub(t, Y1, Y2); fsqr(Dp, t);                                  /* :132 */
  fsub(t, W1p, W2p); fmul(A1p, Y1, t);                           /* :133 */
  fsub(t, Dp, W1p); fsub(X3pc, t, W2p);                          /* :134 */
  fsub(t, X3pc, W1p); fsqr(C, t);                                /* :135 */
  fsub(t, Y1, Y2); fsub(u, W1p, X3pc); fadd(v, t, u); fsqr(w, v);/* :136 */
  fsub(t, w, Dp); fsub(u, t, C); fshl(v, A1p, 1); fsub(Y3p, u, v);
  fshl(t, X3pc, 2); fmul(W1, t, C);                              /* :137 */
  fshl(t, W1p, 2); fmul(W2, t, C);                               /* :138 */
  fshl(t, A1p, 1); fsub(u, Y3p, t); fsqr(D, u);                  /* :139 */
  fsub(t, W1, W2); fmul(A1, Y3p, t);                             /* :140 */
  fsub(t, D, W1); fsub(ret->x, t, W2);                           /* :143 */
  fshl(t, A1p, 1); fsub(u, Y3p, t); fsub(v, W1, ret->x); fmul(w, u, v); fsub(ret->y, w, A1); /* :144 */
  fsub(t, X1, X2); fadd(u, t, X3pc); fsub(v, u, W1p); fsqr(w, v); fsub(t, w, Cp); fsub(u, t, C);
  fmul(ret->z, Z, u);                                            /* :145 */
  fshl(t, A1p, 1); fadd(u, Y3p, t); fsqr(Dc, u);                 /* :147 */
  fsub(t, Dc, W1); fsub(v, t, W2);                               /* :148 (new X2) */
  fshl(t, A1p, 1); fadd(u, Y3p, t); fsub(w, W1, v); fmul(t, u, w); fsub(u, t, A1); /* :149 */
  memcpy(X2, v, 32);
  memcpy(Y2, u, 32);
  memcpy(Q->z, ret->z, 32);                                      /* :150 */
}
/* curve_group.h:155-179 */
static void ADD_Z2_1(jac* ret, const jac* A, const jac* B) {
  const u32 *X1 = A->x, *Y1 = A->y, *Z1 = A->z, *X2 = B->x, *Y2 = B->y;
  u32 Z1Z1[8], U2[8], S2[8], H[8], HH[8], I[8], J[8], r[8], V[8], t[8], u[8], v[8];
  fsqr(Z1Z1, Z1);
  fmul(U2, X2, Z1Z1);
  fmul(t, Y2, Z1); fmul(S2, t, Z1Z1);
  fsub(H, U2, X1);
  fsqr(HH, H);
  fshl(I, HH, 2);
  fmul(J, H, I);
  fsub(t, S2, Y1); fshl(r, t, 1);
  fmul(V, X1, I);
  fsqr(t, r); fsub(u, t, J); fshl(v, V, 1); fsub(ret->x, u, v);
  fsub(t, V, ret->x); fmul(u, r, t); fshl(t, Y1, 1); fmul(v, t, J); fsub(ret->y, u, v);
  fadd(t, Z1, H); fsqr(u, t); fsub(t, u, Z1Z1); fsub(ret->z, t, HH);
}
/* curve_group.h:183-186 */
static void TRPLU(jac* ret, jac* P) {
  jac dbl;
  DBLU(&dbl, P);
  ZADDU(ret, P, &dbl);
}
/* swap.h:47-56: only X and Y are exchanged */
static void swap_xy(jac* A, jac* B) {
  u32 t[8];
  memcpy(t, A->x, 32); memcpy(A->x, B->x, 32); memcpy(B->x, t, 32);
  memcpy(t, A->y, 32); memcpy(A->y, B->y, 32); memcpy(B->y, t, 32);
}
/* curve_group.h:189-218 == lib/scalar_mult_p256.cpp:12-14 */
static void scalar_mult(jac* out, const u32 k[8], const jac* Pin) {
  jac P = *Pin, oppP, base, Psub, nb;
  memcpy(oppP.x, P.x, 32); fopp(oppP.y, P.y); memcpy(oppP.z, P.z, 32); /* jacobian_curve_point.h:48-54 */
  TRPLU(&base, &P);
  if ((k[0] >> 1) & 1) swap_xy(&P, &base);                        /* :192 */
  for (int b = 2; b < 256; b++) {                                  /* :194-213 */
    int bit = (k[b >> 5] >> (b & 31)) & 1;
    if (bit) swap_xy(&P, &base);
    ZDAU(&nb, &base, &P);
    base = nb;
    if (bit) swap_xy(&P, &base);
  }
  ADD_Z2_1(&Psub, &P, &oppP);                                      /* :216 */
  *out = (k[0] & 1) ? P : Psub;                                    /* :215,217 */
}
/* jacobian_curve_point.h:25-31 */
static void from_affine(jac* J, const u32 xy[16]) {
  from_classical(J->x, xy);
  from_classical(J->y, xy + 8);
  memcpy(J->z, R_P, 32);
}
/* jacobian_curve_point.h:33-42 */
static void to_affine(u32 xy[16], const jac* J) {
  u32 invZ[8], invZ2[8], invZ3[8], t[8];
  finv(invZ, J->z);
  fsqr(invZ2, invZ);
  fmul(invZ3, invZ2, invZ);
  fmul(t, J->x, invZ2); to_classical(xy, t);
  fmul(t, J->y, invZ3); to_classical(xy + 8, t);
}
/* curve_group.h:43-58 */
static int from_x(u32 y[8], const u32 x[8]) {
  u32 xm[8], xpow3[8], x3[8], ypow2[8], t[8], u[8], ym[8];
  from_classical(xm, x);
  fsqr(t, xm); fmul(xpow3, t, xm);
  fshl(t, xm, 1); fadd(x3, t, xm);
  fadd(u, xpow3, BM); fsub(ypow2, u, x3);
  int ok = fsqrt(ym, ypow2);
  to_classical(y, ym);
  return ok;
}

/* ---- batched drivers -------------------------------------------------------- */
typedef void (*range_fn)(void* ctx, size_t lo, size_t hi);
typedef struct { range_fn f; void* ctx; size_t lo, hi; } job_t;
static void* job_main(void* p) { job_t* j = (job_t*)p; j->f(j->ctx, j->lo, j->hi); return NULL; }
static void par_range(size_t n, int nt, range_fn f, void* ctx) {
  if (nt <= 1 || n < 2) { f(ctx, 0, n); return; }
  if (nt > 256) nt = 256;
  pthread_t th[256];
  job_t jobs[256];
  size_t per = (n + (size_t)nt - 1) / (size_t)nt;
  int started = 0;
  for (int t = 0; t < nt; t++) {
    size_t lo = (size_t)t * per, hi = lo + per > n ? n : lo + per;
    if (lo >= hi) break;
    jobs[t] = (job_t){f, ctx, lo, hi};
    pthread_create(&th[t], NULL, job_main, &jobs[t]);
    started++;
  }
  for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
}

typedef struct { u32 *o, *o2; const u32 *a, *b; uint8_t* ok; } args_t;

#define DEF_RANGE(name, ...)                                       \
  static void name(void* c, size_t lo, size_t hi) {                \
    args_t* g = (args_t*)c;                                        \
    for (size_t i = lo; i < hi; i++) { __VA_ARGS__; }              \
  }
DEF_RANGE(r_add, orc1_mod_add(g->o + 8 * i, g->a + 8 * i, g->b + 8 * i))
DEF_RANGE(r_sub, orc1_mod_sub(g->o + 8 * i, g->a + 8 * i, g->b + 8 * i))
DEF_RANGE(r_shl, orc1_mod_shl1(g->o + 8 * i, g->a + 8 * i))
DEF_RANGE(r_mul, fmul(g->o + 8 * i, g->a + 8 * i, g->b + 8 * i))
DEF_RANGE(r_sqr, fsqr(g->o + 8 * i, g->a + 8 * i))
DEF_RANGE(r_opp, fopp(g->o + 8 * i, g->a + 8 * i))
DEF_RANGE(r_fc, from_classical(g->o + 8 * i, g->a + 8 * i))
DEF_RANGE(r_tc, to_classical(g->o + 8 * i, g->a + 8 * i))
DEF_RANGE(r_inv, finv(g->o + 8 * i, g->a + 8 * i))
DEF_RANGE(r_dblu, { jac p, r; memcpy(&p, g->a + 24 * i, 96); DBLU(&r, &p);
                    memcpy(g->o + 24 * i, &p, 96); memcpy(g->o2 + 24 * i, &r, 96); })
DEF_RANGE(r_zaddu, { jac p, o, r; memcpy(&p, g->a + 24 * i, 96); memcpy(&o, g->b + 24 * i, 96); ZADDU(&r, &p, &o);
                     memcpy(g->o + 24 * i, &p, 96); memcpy(g->o2 + 24 * i, &r, 96); })
DEF_RANGE(r_zdau, { jac p, q, r; memcpy(&p, g->a + 24 * i, 96); memcpy(&q, g->b + 24 * i, 96); ZDAU(&r, &p, &q);
                    memcpy(g->o + 24 * i, &q, 96); memcpy(g->o2 + 24 * i, &r, 96); })
DEF_RANGE(r_addz, { jac a, b, r; memcpy(&a, g->a + 24 * i, 96); memcpy(&b, g->b + 24 * i, 96); ADD_Z2_1(&r, &a, &b);
                    memcpy(g->o + 24 * i, &r, 96); })
DEF_RANGE(r_trplu, { jac p, r; memcpy(&p, g->a + 24 * i, 96); TRPLU(&r, &p);
                     memcpy(g->o + 24 * i, &p, 96); memcpy(g->o2 + 24 * i, &r, 96); })
DEF_RANGE(r_smul, { jac p, r; memcpy(&p, g->b + 24 * i, 96); scalar_mult(&r, g->a + 8 * i, &p);
                    memcpy(g->o + 24 * i, &r, 96); })
typedef struct { u32* o; const u32 *k, *P; u32* wraps; } wargs_t;
static void r_smul_wraps(void* c, size_t lo, size_t hi) {
  wargs_t* g = (wargs_t*)c;
  for (size_t i = lo; i < hi; i++) {
    jac p, r;
    u64 before = g_cnt[5];
    memcpy(&p, g->P + 24 * i, 96);
    scalar_mult(&r, g->k + 8 * i, &p);
    memcpy(g->o + 24 * i, &r, 96);
    g->wraps[i] = (u32)(g_cnt[5] - before);
  }
}
DEF_RANGE(r_fa, { jac r; from_affine(&r, g->a + 16 * i); memcpy(g->o + 24 * i, &r, 96); })
DEF_RANGE(r_ta, { jac p; memcpy(&p, g->a + 24 * i, 96); to_affine(g->o + 16 * i, &p); })
DEF_RANGE(r_fx, g->ok[i] = (uint8_t)from_x(g->o + 8 * i, g->a + 8 * i))

#define RUN(fn, O, O2, A, B, OK) do { args_t g = {(O), (O2), (A), (B), (OK)}; par_range(n, nt, fn, &g); } while (0)

void orc_mgry_add(u32* o, const u32* a, const u32* b, size_t n, int nt) { RUN(r_add, o, 0, a, b, 0); }
void orc_mgry_sub(u32* o, const u32* a, const u32* b, size_t n, int nt) { RUN(r_sub, o, 0, a, b, 0); }
void orc_mgry_shl1(u32* o, const u32* a, size_t n, int nt) { RUN(r_shl, o, 0, a, 0, 0); }
void orc_mgry_mul(u32* o, const u32* a, const u32* b, size_t n, int nt) { RUN(r_mul, o, 0, a, b, 0); }
void orc_mgry_sqr(u32* o, const u32* a, size_t n, int nt) { RUN(r_sqr, o, 0, a, 0, 0); }
void orc_opposite(u32* o, const u32* a, size_t n, int nt) { RUN(r_opp, o, 0, a, 0, 0); }
void orc_from_classical(u32* o, const u32* a, size_t n, int nt) { RUN(r_fc, o, 0, a, 0, 0); }
void orc_to_classical(u32* o, const u32* a, size_t n, int nt) { RUN(r_tc, o, 0, a, 0, 0); }
void orc_inverse(u32* o, const u32* a, size_t n, int nt) { RUN(r_inv, o, 0, a, 0, 0); }
void orc_mul512(u32* o, const u32* a, const u32* b, size_t n) { for (size_t i = 0; i < n; i++) orc1_mul512(o + 16 * i, a + 8 * i, b + 8 * i); }
void orc_square512(u32* o, const u32* a, size_t n) { for (size_t i = 0; i < n; i++) orc1_square512(o + 16 * i, a + 8 * i); }
void orc_mgry_reduce(u32* o, const u32* t, size_t n) { for (size_t i = 0; i < n; i++) orc1_mgry_reduce(o + 8 * i, t + 16 * i); }
void orc_dblu(u32* outP, u32* out2, const u32* P, size_t n, int nt) { RUN(r_dblu, outP, out2, P, 0, 0); }
void orc_zaddu(u32* outP, u32* outR, const u32* P, const u32* O, size_t n, int nt) { RUN(r_zaddu, outP, outR, P, O, 0); }
void orc_zdau(u32* outQ, u32* outR, const u32* P, const u32* Q, size_t n, int nt) { RUN(r_zdau, outQ, outR, P, Q, 0); }
void orc_add_z2_1(u32* outR, const u32* A, const u32* B, size_t n, int nt) { RUN(r_addz, outR, 0, A, B, 0); }
void orc_trplu(u32* outP, u32* out3, const u32* P, size_t n, int nt) { RUN(r_trplu, outP, out3, P, 0, 0); }
void orc_scalar_mult(u32* out, const u32* k, const u32* P, size_t n, int nt) { RUN(r_smul, out, 0, k, P, 0); }
/* scalar_mult that also reports, per lane, how many times square() lost a carry */
void orc_scalar_mult_wraps(u32* out, u32* wraps, const u32* k, const u32* P, size_t n, int nt) {
  wargs_t g = {out, k, P, wraps};
  par_range(n, nt, r_smul_wraps, &g);
}
void orc_from_affine(u32* outJ, const u32* xy, size_t n, int nt) { RUN(r_fa, outJ, 0, xy, 0, 0); }
void orc_to_affine(u32* xy, const u32* J, size_t n, int nt) { RUN(r_ta, xy, 0, J, 0, 0); }
void orc_from_x(u32* y, uint8_t* ok, const u32* x, size_t n, int nt) { RUN(r_fx, y, 0, x, 0, ok); }

/* ---- the same field layer for a modulus given at run time ---------------------------------------
 * (the reference's templates take the prime as a parameter; its field tests use secp256k1:
 * tests/mgry.cpp:25-27, tests/ops.cpp:221-252).  p: 8 words, odd, bit 255 set. */
typedef struct { u32 p[8], r1[8], rr[8], mprime; } genp;
static u32 geq8(const u32* a, const u32* p) { for (int i = 7; i >= 0; i--) if (a[i] != p[i]) return a[i] > p[i]; return 1; }
static void gen_setup(genp* g, const u32* p) {
  memcpy(g->p, p, 32);
  u32 inv = p[0];
  for (int i = 0; i < 5; i++) inv *= 2u - p[0] * inv;
  g->mprime = 0u - inv;                                   /* mgry_mul.h:33-40 */
  u32 x[8] = {1, 0, 0, 0, 0, 0, 0, 0}, d[8];
  for (int i = 0; i < 512; i++) {                          /* R mod p, R^2 mod p: mgry_csts.h:20-21 */
    u32 c = x[7] >> 31;
    for (int k = 7; k > 0; k--) x[k] = (x[k] << 1) | (x[k - 1] >> 31);
    x[0] <<= 1;
    if (c || geq8(x, p)) { sub8(d, x, p); memcpy(x, d, 32); }
    if (i == 255) memcpy(g->r1, x, 32);
  }
  memcpy(g->rr, x, 32);
}
static void g_reduce_once(u32 r[8], const u32 s[8], u32 c, const genp* g) {  /* sub.h:46-69 */
  u32 d[8];
  u32 bw = sub8(d, s, g->p);
  memcpy(r, (bw && !c) ? s : d, 32);
}
static void g_add(u32 r[8], const u32 a[8], const u32 b[8], const genp* g) { u32 s[8]; u32 c = add8(s, a, b); g_reduce_once(r, s, c, g); }
static void g_sub(u32 r[8], const u32 a[8], const u32 b[8], const genp* g) {
  u32 d[8], da[8];
  u32 bw = sub8(d, a, b);
  add8(da, d, g->p);
  memcpy(r, bw ? da : d, 32);
}
static void g_shl1(u32 r[8], const u32 a[8], const genp* g) {
  u32 s[8];
  u32 c = a[7] >> 31;
  for (int i = 7; i > 0; i--) s[i] = (a[i] << 1) | (a[i - 1] >> 31);
  s[0] = a[0] << 1;
  g_reduce_once(r, s, c, g);
}
static void g_redc(u32 r[8], const u32 t[16], const genp* g) {  /* mgry_mul.h:84-121, generic digits of p */
  u64 acc[17];
  for (int k = 0; k < 16; k++) acc[k] = t[k];
  acc[16] = 0;
  for (int i = 0; i < 8; i++) {
    u32 m = (u32)acc[i] * g->mprime;
    u64 carry = 0;
    for (int k = 0; k < 8; k++) { u64 x = acc[i + k] + (u64)m * g->p[k] + carry; acc[i + k] = x & M32; carry = x >> 32; }
    for (int k = i + 8; k < 17; k++) { u64 x = acc[k] + carry; acc[k] = x & M32; carry = x >> 32; }
  }
  u32 s[8];
  for (int k = 0; k < 8; k++) s[k] = (u32)acc[8 + k];
  g_reduce_once(r, s, (u32)acc[16], g);
}
static void g_mul(u32 r[8], const u32 a[8], const u32 b[8], const genp* g) { u32 t[16]; orc1_mul512(t, a, b); g_redc(r, t, g); }
static void g_sqr(u32 r[8], const u32 a[8], const genp* g) { u32 t[16]; square512_impl(t, a); g_redc(r, t, g); }
static void g_pow(u32 r[8], const u32 a[8], const u32 e[8], const genp* g) {  /* mgry_ops.h:44-86 */
  int top = -1;
  for (int b = 255; b >= 0; b--) if ((e[b >> 5] >> (b & 31)) & 1) { top = b; break; }
  u32 res[8], base[8], t[8];
  memcpy(res, g->r1, 32);
  memcpy(base, a, 32);
  for (int b = 0; b <= top; b++) {
    if ((e[b >> 5] >> (b & 31)) & 1) { g_mul(t, res, base, g); memcpy(res, t, 32); }
    if (b < top) { g_sqr(t, base, g); memcpy(base, t, 32); }
  }
  memcpy(r, res, 32);
}
/* op: 0 mod_add 1 mod_sub 2 mod_shift_left_one 3 mgry_mul 4 mgry_sqr 5 from_classical 6 to_classical
 *     7 mgry_pow (e) 8 opposite */
void orc_gen_op(int op, u32* o, const u32* a, const u32* b, const u32* e, const u32* p, size_t n) {
  genp g;
  gen_setup(&g, p);
  u32 one[8] = {1, 0, 0, 0, 0, 0, 0, 0}, pm1r[8], t[8];
  sub8(pm1r, g.p, g.r1);                                   /* (p-1)R mod p = p - (R mod p) */
  for (size_t i = 0; i < n; i++) {
    const u32 *x = a + 8 * i, *y = b ? b + 8 * i : NULL;
    u32* r = o + 8 * i;
    switch (op) {
      case 0: g_add(r, x, y, &g); break;
      case 1: g_sub(r, x, y, &g); break;
      case 2: g_shl1(r, x, &g); break;
      case 3: g_mul(r, x, y, &g); break;
      case 4: g_sqr(r, x, &g); break;
      case 5: g_mul(r, x, g.rr, &g); break;
      case 6: g_mul(r, x, one, &g); break;
      case 7: g_pow(r, x, e, &g); break;
      default: g_sub(t, x, g.r1, &g); g_sub(r, pm1r, t, &g); break;   /* gfp.h:60-64 */
    }
  }
}

void orc_constants(u32* out) {
  memcpy(out, P256, 32); memcpy(out + 8, R_P, 32); memcpy(out + 16, RSQ_P, 32); memcpy(out + 24, PM1_R_P, 32);
  memcpy(out + 32, AM, 32); memcpy(out + 40, BM, 32);
  jac g; u32 xy[16]; memcpy(xy, GX, 32); memcpy(xy + 8, GY, 32); from_affine(&g, xy);
  memcpy(out + 48, g.x, 32); memcpy(out + 56, g.y, 32);
}
