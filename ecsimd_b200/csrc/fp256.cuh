// fp256.cuh -- GF(p256) arithmetic in Montgomery form (R = 2^256) on 8 x 32-bit
// limbs held in registers, for sm_100a.
//
// Every function reproduces, bit for bit and for ANY 256-bit input pattern, the
// value the reference computes per lane (paths relative to aguinet/ecsimd):
//   fp_add   = mgry_add          include/ecsimd/mgry_ops.h:10-13  (modular.h:10-15, sub.h:46-69)
//   fp_sub   = mgry_sub          include/ecsimd/mgry_ops.h:26-29  (modular.h:24-41)
//   fp_shl1  = mgry_shift_left<1> include/ecsimd/mgry_ops.h:15-24 (modular.h:17-22, shift.h:13-32)
//   fp_mul   = mgry_mul          include/ecsimd/mgry_ops.h:31-35  (mul.h:150-158, mgry_mul.h:84-121)
//   fp_sqr   = mgry_sqr          include/ecsimd/mgry_ops.h:37-42  (mul.h:160-221 -- including its
//                                 lost-carry defect, see fp_sqr below)
//   fp_neg   = GFp::opposite     include/ecsimd/gfp.h:60-64
// The implementation is not a translation of the AVX2 code: see gen_fp256.py
// for the multiplier design (IMAD.WIDE carry chains, 3-term P-256 reduction).
#pragma once
#include <cstdint>

#include "fp256_mul_gen.cuh"

namespace ecb200 {

struct fe {
  uint32_t v[8];
};

// p, R mod p, R^2 mod p, (p-1)R mod p, Am = -3R, Bm = bR (least-significant word first).
// mgry_csts.h:20-24, curve_group.h:31-32 of the reference; values cross-checked by
// tests/test_oracle_vs_ref.py::test_constants.
#define ECB200_P_WORDS   {0xffffffffu, 0xffffffffu, 0xffffffffu, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000001u, 0xffffffffu}
#define ECB200_R_WORDS   {0x00000001u, 0x00000000u, 0x00000000u, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xfffffffeu, 0x00000000u}
#define ECB200_RR_WORDS  {0x00000003u, 0x00000000u, 0xffffffffu, 0xfffffffbu, 0xfffffffeu, 0xffffffffu, 0xfffffffdu, 0x00000004u}
#define ECB200_PM1R_WORDS {0xfffffffeu, 0xffffffffu, 0xffffffffu, 0x00000001u, 0x00000000u, 0x00000000u, 0x00000002u, 0xfffffffeu}
#define ECB200_AM_WORDS  {0xfffffffcu, 0xffffffffu, 0xffffffffu, 0x00000003u, 0x00000000u, 0x00000000u, 0x00000004u, 0xfffffffcu}
#define ECB200_BM_WORDS  {0x29c4bddfu, 0xd89cdf62u, 0x78843090u, 0xacf005cdu, 0xf7212ed6u, 0xe5a220abu, 0x04874834u, 0xdc30061du}
// G in Montgomery form (from_affine(Gx,Gy)), jacobian_curve_point.h:25-31 on curve_nist_p256.h:26-31
#define ECB200_GXM_WORDS {0x18a9143cu, 0x79e730d4u, 0x5fedb601u, 0x75ba95fcu, 0x77622510u, 0x79fb732bu, 0xa53755c6u, 0x18905f76u}
#define ECB200_GYM_WORDS {0xce95560au, 0xddf25357u, 0xba19e45cu, 0x8b4ab8e4u, 0xdd21f325u, 0xd2e88688u, 0x25885d85u, 0x8571ff18u}

__device__ __forceinline__ fe fe_const(const uint32_t (&w)[8]) {
  fe r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = w[i];
  return r;
}
__device__ __forceinline__ fe fe_R() { const uint32_t w[8] = ECB200_R_WORDS; return fe_const(w); }
__device__ __forceinline__ fe fe_RR() { const uint32_t w[8] = ECB200_RR_WORDS; return fe_const(w); }
__device__ __forceinline__ fe fe_PM1R() { const uint32_t w[8] = ECB200_PM1R_WORDS; return fe_const(w); }
__device__ __forceinline__ fe fe_AM() { const uint32_t w[8] = ECB200_AM_WORDS; return fe_const(w); }
__device__ __forceinline__ fe fe_BM() { const uint32_t w[8] = ECB200_BM_WORDS; return fe_const(w); }
__device__ __forceinline__ fe fe_zero() { fe r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = 0;
  return r; }

// ---- a - b, plus p when the subtraction borrowed ---------------------------------
// p & mask = {M, M, M, 0, 0, 0, M&1, M}
__device__ __forceinline__ fe fp_sub(const fe& a, const fe& b) {
  fe d;
  uint32_t mk;
  asm("sub.cc.u32 %0, %9, %17; subc.cc.u32 %1, %10, %18; subc.cc.u32 %2, %11, %19; subc.cc.u32 %3, %12, %20; "
      "subc.cc.u32 %4, %13, %21; subc.cc.u32 %5, %14, %22; subc.cc.u32 %6, %15, %23; subc.cc.u32 %7, %16, %24; "
      "subc.u32 %8, 0, 0;"
      : "=r"(d.v[0]), "=r"(d.v[1]), "=r"(d.v[2]), "=r"(d.v[3]), "=r"(d.v[4]), "=r"(d.v[5]), "=r"(d.v[6]), "=r"(d.v[7]), "=r"(mk)
      : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
        "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
  uint32_t one = mk & 1u;
  asm("add.cc.u32 %0, %0, %8; addc.cc.u32 %1, %1, %8; addc.cc.u32 %2, %2, %8; addc.cc.u32 %3, %3, 0; "
      "addc.cc.u32 %4, %4, 0; addc.cc.u32 %5, %5, 0; addc.cc.u32 %6, %6, %9; addc.u32 %7, %7, %8;"
      : "+r"(d.v[0]), "+r"(d.v[1]), "+r"(d.v[2]), "+r"(d.v[3]), "+r"(d.v[4]), "+r"(d.v[5]), "+r"(d.v[6]), "+r"(d.v[7])
      : "r"(mk), "r"(one));
  return d;
}

// t_lo >= p, for the rare case t7 == 0xffffffff (p = ffffffff 00000001 00000000 00000000 00000000 ffffffff ffffffff ffffffff)
static __device__ __noinline__ uint32_t fp_ge_p_rare(uint32_t t0, uint32_t t1, uint32_t t2, uint32_t t3, uint32_t t4, uint32_t t5, uint32_t t6) {
  if (t6 != 1u) return t6 > 1u;
  if ((t3 | t4 | t5) != 0u) return 1u;
  return (t0 & t1 & t2) == 0xffffffffu;
}

// Two ways of treating the (astronomically) rare cases of the fast paths below:
//   Exact : resolve them in place with a branch (streaming field / point kernels).
//   Lazy  : only record that one may have happened; the caller re-runs the whole computation
//           in Exact mode if anything was recorded (the scalar-mult ladder: keeps branches and
//           calls out of its hot loop).  `top` collects the largest pre-reduction top word seen
//           (0xffffffff <=> the conditional subtraction may have been decided wrongly);
//           A squaring operand in the reference's lost-carry set is repaired in place in both modes
//           (the exact test is already behind a rarely taken branch; `dirty` is kept for callers
//           that want to force the exact re-run).
struct Exact {};
struct Lazy {
  uint32_t top = 0;
  uint32_t dirty = 0;
  __device__ __forceinline__ bool flagged() const { return top == 0xffffffffu || dirty != 0u; }
};

// The single conditional subtraction every modular op ends with (sub_if_above, sub.h:46-69):
// given the 257-bit value (c:s), return it minus p if it is >= p.  c == 1 always subtracts;
// c == 0 subtracts only if s >= p, which needs s7 == 0xffffffff (probability 2^-32 on uniform
// data): that case branches to an exact comparison (or is only flagged, Lazy mode); everything else
// is one 8-word subtraction of p & mask.
__device__ __forceinline__ fe fp_add_k(const fe& s, uint32_t sub) {
  // s - (p & mask), p & mask = {mask, mask, mask, 0, 0, 0, sub, mask}   (mod 2^256: the borrow out cancels c)
  const uint32_t mask = 0u - sub;
  fe r;
  asm("sub.cc.u32 %0, %8, %17; subc.cc.u32 %1, %9, %17; subc.cc.u32 %2, %10, %17; subc.cc.u32 %3, %11, 0; "
      "subc.cc.u32 %4, %12, 0; subc.cc.u32 %5, %13, 0; subc.cc.u32 %6, %14, %16; subc.u32 %7, %15, %17;"
      : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7])
      : "r"(s.v[0]), "r"(s.v[1]), "r"(s.v[2]), "r"(s.v[3]), "r"(s.v[4]), "r"(s.v[5]), "r"(s.v[6]), "r"(s.v[7]),
        "r"(sub), "r"(mask));
  return r;
}
__device__ __forceinline__ fe fp_reduce_once(const fe& s, uint32_t c, Exact&) {
  uint32_t sub = c;
  if (__builtin_expect(s.v[7] == 0xffffffffu && c == 0u, 0))
    sub = fp_ge_p_rare(s.v[0], s.v[1], s.v[2], s.v[3], s.v[4], s.v[5], s.v[6]);
  return fp_add_k(s, sub);
}
__device__ __forceinline__ fe fp_reduce_once(const fe& s, uint32_t c, Lazy& z) {
  z.top = max(z.top, s.v[7]);
  return fp_add_k(s, c);
}
__device__ __forceinline__ fe fp_reduce_once(const fe& s, uint32_t c) { Exact e; return fp_reduce_once(s, c, e); }

template <class M>
__device__ __forceinline__ fe fp_add(const fe& a, const fe& b, M& mode) {
  fe s;
  uint32_t c;
  asm("add.cc.u32 %0, %9, %17; addc.cc.u32 %1, %10, %18; addc.cc.u32 %2, %11, %19; addc.cc.u32 %3, %12, %20; "
      "addc.cc.u32 %4, %13, %21; addc.cc.u32 %5, %14, %22; addc.cc.u32 %6, %15, %23; addc.cc.u32 %7, %16, %24; "
      "addc.u32 %8, 0, 0;"
      : "=r"(s.v[0]), "=r"(s.v[1]), "=r"(s.v[2]), "=r"(s.v[3]), "=r"(s.v[4]), "=r"(s.v[5]), "=r"(s.v[6]), "=r"(s.v[7]), "=r"(c)
      : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
        "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
  return fp_reduce_once(s, c, mode);
}
__device__ __forceinline__ fe fp_add(const fe& a, const fe& b) { Exact e; return fp_add(a, b, e); }

template <class M>
__device__ __forceinline__ fe fp_shl1(const fe& a, M& mode) {
  fe s;
  const uint32_t c = a.v[7] >> 31;
#pragma unroll
  for (int i = 7; i > 0; i--) s.v[i] = __funnelshift_l(a.v[i - 1], a.v[i], 1);
  s.v[0] = a.v[0] << 1;
  return fp_reduce_once(s, c, mode);
}
__device__ __forceinline__ fe fp_shl1(const fe& a) { Exact e; return fp_shl1(a, e); }
template <int COUNT, class M>
__device__ __forceinline__ fe fp_shl(const fe& a, M& mode) {
  fe r = a;
#pragma unroll
  for (int i = 0; i < COUNT; i++) r = fp_shl1(r, mode);
  return r;
}
template <int COUNT>
__device__ __forceinline__ fe fp_shl(const fe& a) { Exact e; return fp_shl<COUNT>(a, e); }

// 4a mod p for a value that only feeds multiplications: any representative below 2^256 will do
// there (fp_mul takes ANY 256-bit pattern and returns the canonical residue), so one pass replaces
// two canonical doublings.  4a = s + t*2^256 with s = (a << 2) mod 2^256, t = a >> 254, and
// 4a - t*p = s + t*(2^256 - p); t*(2^256 - p) = {t, 0, 0, -t, m, m, ~t & m, t + m}, m = -(t != 0).
// The sum can pass 2^256 only if s >= 2^256 - 3*2^224: recorded in z.top like every other rare case.
__device__ __forceinline__ fe fp_shl2_mulonly(const fe& a, Lazy& z) {
  fe s;
#pragma unroll
  for (int i = 7; i > 0; i--) s.v[i] = __funnelshift_l(a.v[i - 1], a.v[i], 2);
  s.v[0] = a.v[0] << 2;
  const uint32_t t = a.v[7] >> 30;
  const uint32_t nt = 0u - t;
  const uint32_t m = (uint32_t)((int32_t)nt >> 31);
  const uint32_t w6 = m & ~t, w7 = t + m;
  z.top = max(z.top, s.v[7] | 3u);
  fe r;
  asm("add.cc.u32 %0, %8, %16; addc.cc.u32 %1, %9, 0; addc.cc.u32 %2, %10, 0; addc.cc.u32 %3, %11, %17; "
      "addc.cc.u32 %4, %12, %18; addc.cc.u32 %5, %13, %18; addc.cc.u32 %6, %14, %19; addc.u32 %7, %15, %20;"
      : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7])
      : "r"(s.v[0]), "r"(s.v[1]), "r"(s.v[2]), "r"(s.v[3]), "r"(s.v[4]), "r"(s.v[5]), "r"(s.v[6]), "r"(s.v[7]),
        "r"(t), "r"(nt), "r"(m), "r"(w6), "r"(w7));
  return r;
}
__device__ __forceinline__ fe fp_shl2_mulonly(const fe& a, Exact& e) { return fp_shl<2>(a, e); }

template <class M>
__device__ __forceinline__ fe fp_mul(const fe& a, const fe& b, M& mode) {
  fe t;
  uint32_t t8;
  fp_mul_t9(t.v[0], t.v[1], t.v[2], t.v[3], t.v[4], t.v[5], t.v[6], t.v[7], t8,
            a.v[0], a.v[1], a.v[2], a.v[3], a.v[4], a.v[5], a.v[6], a.v[7],
            b.v[0], b.v[1], b.v[2], b.v[3], b.v[4], b.v[5], b.v[6], b.v[7]);
  return fp_reduce_once(t, t8, mode);
}
__device__ __forceinline__ fe fp_mul(const fe& a, const fe& b) { Exact e; return fp_mul(a, b, e); }

// ---- a*b + c with one reduction ---------------------------------------------------------
// fp_mul_wide: the unreduced 512-bit product.  fp_mul_acc(c, a, b) = (a*b + c) * R^-1 mod p, canonical:
// the value mgry_add(mgry_mul(a, b), mgry_reduce(c)) of the reference, for one reduction instead of
// two and no separate addition.  The quotient t = (a*b + c + m*p) / 2^256 is below 2^257 + p, i.e.
// t = (k : s) with k in {0, 1, 2}, and t - k*p = s + k*(2^256 - p) is below 2^256 and canonical
// unless s is within 2^226 of a multiple of 2^256 (k*(2^256 - p) = {k, 0, 0, -k, m, m, m & ~k, k + m},
// m = -(k != 0)): the fast path is one 8-word add; the rare cases show as an all-ones top word before
// or after it.
struct fe512 {
  uint32_t v[16];
};
__device__ __forceinline__ fe512 fp_mul_wide(const fe& a, const fe& b) {
  fe512 t;
  fp_mul512_words(t.v[0], t.v[1], t.v[2], t.v[3], t.v[4], t.v[5], t.v[6], t.v[7], t.v[8], t.v[9], t.v[10], t.v[11], t.v[12], t.v[13],
                  t.v[14], t.v[15], a.v[0], a.v[1], a.v[2], a.v[3], a.v[4], a.v[5], a.v[6], a.v[7],
                  b.v[0], b.v[1], b.v[2], b.v[3], b.v[4], b.v[5], b.v[6], b.v[7]);
  return t;
}
__device__ __forceinline__ void fp_mul_acc_quot(fe& s, uint32_t& k, const fe512& c, const fe& a, const fe& b) {
  fp_mul_acc_t9(s.v[0], s.v[1], s.v[2], s.v[3], s.v[4], s.v[5], s.v[6], s.v[7], k,
                a.v[0], a.v[1], a.v[2], a.v[3], a.v[4], a.v[5], a.v[6], a.v[7],
                b.v[0], b.v[1], b.v[2], b.v[3], b.v[4], b.v[5], b.v[6], b.v[7],
                c.v[0], c.v[1], c.v[2], c.v[3], c.v[4], c.v[5], c.v[6], c.v[7], c.v[8], c.v[9], c.v[10], c.v[11], c.v[12], c.v[13],
                c.v[14], c.v[15]);
}
__device__ __forceinline__ fe fp_mul_acc(const fe512& c, const fe& a, const fe& b, Lazy& z) {
  fe s;
  uint32_t k;
  fp_mul_acc_quot(s, k, c, a, b);
  const uint32_t nk = 0u - k;
  const uint32_t m = (uint32_t)((int32_t)nk >> 31);
  const uint32_t w6 = m & ~k, w7 = k + m;
  fe r;
  asm("add.cc.u32 %0, %8, %16; addc.cc.u32 %1, %9, 0; addc.cc.u32 %2, %10, 0; addc.cc.u32 %3, %11, %17; "
      "addc.cc.u32 %4, %12, %18; addc.cc.u32 %5, %13, %18; addc.cc.u32 %6, %14, %19; addc.u32 %7, %15, %20;"
      : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7])
      : "r"(s.v[0]), "r"(s.v[1]), "r"(s.v[2]), "r"(s.v[3]), "r"(s.v[4]), "r"(s.v[5]), "r"(s.v[6]), "r"(s.v[7]),
        "r"(k), "r"(nk), "r"(m), "r"(w6), "r"(w7));
  z.top = __vimax3_u32(z.top, s.v[7], r.v[7]);
  return r;
}
// exact form (the flagged lanes' re-run): subtract p while the 258-bit value (k : s) is >= p
static __device__ __noinline__ void fp_sub_p_while_ge(uint32_t* s, uint32_t k) {
  const uint32_t pw[8] = ECB200_P_WORDS;
  for (int it = 0; it < 3; it++) {
    bool ge = k != 0u;
    if (!ge) {
      ge = true;  // s >= p ?
      for (int i = 7; i >= 0; i--) {
        if (s[i] != pw[i]) { ge = s[i] > pw[i]; break; }
      }
    }
    if (!ge) return;
    uint32_t borrow = 0;
    for (int i = 0; i < 8; i++) {
      const unsigned long long d = (unsigned long long)s[i] - pw[i] - borrow;
      s[i] = (uint32_t)d;
      borrow = (uint32_t)(d >> 63);
    }
    k -= borrow;
  }
}
__device__ __forceinline__ fe fp_mul_acc(const fe512& c, const fe& a, const fe& b, Exact&) {
  fe s;
  uint32_t k;
  fp_mul_acc_quot(s, k, c, a, b);
  uint32_t w[8];
#pragma unroll
  for (int i = 0; i < 8; i++) w[i] = s.v[i];
  fp_sub_p_while_ge(w, k);
#pragma unroll
  for (int i = 0; i < 8; i++) s.v[i] = w[i];
  return s;
}

// ---- squaring ------------------------------------------------------------------------
// The reference's square() (mul.h:160-212) accumulates the doubled cross products
// `2*a_i*a_j + ret + prev` in wrap-around 64-bit lanes (mul.h:192-195); when that
// sum reaches 2^64 a carry is silently lost, so mgry_sqr(a) != mgry_mul(a,a) for
// about 2e-9 of random inputs.  Bit-exact parity needs the same answer.
//
// A wrap needs 2*pr mod 2^64 >= 2^64 - 2^33 - 1 for some cross product
// pr = a_i*a_j (the other two addends are < 2^32+2 and <= 2^32), i.e. bit 63 of
// pr clear and bits 62..32 all set: the high word of pr, read as a signed int,
// is INT_MAX.  So: fast path = true square (fp_sqr_t9) plus a cheap necessary-condition
// filter (fp_sqr_quirk_filter below); only lanes that pass it AND the exact test
// (probability ~ 6.5e-9 per lane) re-run the reference's loop literally (fp_sqr_quirk_slow).
static __device__ __noinline__ void fp_sqr_quirk_slow(uint32_t* r, const uint32_t* a) {
  // literal restatement of mul.h:176-210 on 64-bit wrap-around integers
  unsigned long long ret[17];
  for (int k = 0; k < 17; k++) ret[k] = 0;
  for (int i = 0; i < 8; i++) {
    unsigned long long t = (unsigned long long)a[i] * a[i] + ret[2 * i];
    ret[2 * i] = t & 0xffffffffull;
    unsigned long long p0 = t >> 32, p1 = 0;
    for (int j = i + 1; j < 8; j++) {
      unsigned long long pr = (unsigned long long)a[i] * a[j];
      unsigned long long carry = pr >> 63;
      t = (pr << 1) + ret[i + j] + p0;  // may wrap: that is the defect being reproduced
      ret[i + j] = t & 0xffffffffull;
      p0 = p1 + (t >> 32);
      p1 = carry;
    }
    ret[i + 8] += p0;
    if (i + 9 < 16) ret[i + 9] = p1;
  }
  // trunc_u64x32 (mul.h:85-113) then mgry_reduce (mgry_mul.h:84-121), m' = 1
  unsigned long long acc[17];
  for (int k = 0; k < 16; k++) acc[k] = ret[k] & 0xffffffffull;
  acc[16] = 0;
  const uint32_t pw[8] = ECB200_P_WORDS;
  for (int i = 0; i < 8; i++) {
    const unsigned long long m = acc[i];
    unsigned long long carry = 0;
    for (int k = 0; k < 8; k++) {
      unsigned long long x = acc[i + k] + m * pw[k] + carry;
      acc[i + k] = x & 0xffffffffull;
      carry = x >> 32;
    }
    for (int k = i + 8; k < 17 && carry; k++) {
      unsigned long long x = acc[k] + carry;
      acc[k] = x & 0xffffffffull;
      carry = x >> 32;
    }
  }
  long long bw = 0;
  uint32_t d[8];
  for (int k = 0; k < 8; k++) {
    long long x = (long long)acc[8 + k] - (long long)pw[k] - bw;
    d[k] = (uint32_t)x;
    bw = (x < 0) ? 1 : 0;
  }
  const bool lt = ((long long)acc[16] - bw) < 0;  // t < p
  for (int k = 0; k < 8; k++) r[k] = lt ? (uint32_t)acc[8 + k] : d[k];
}

// exact test: is some cross-product high word 0x7fffffff (INT_MAX as a signed int)?
static __device__ __noinline__ uint32_t fp_sqr_quirk_filter_exact(const uint32_t* a) {
  // straight-line: this runs for ~1 lane in 1 200 squarings (first-level false positives), i.e. in
  // about one ladder step in six per warp, so its length matters a little
  uint32_t w[8];
#pragma unroll
  for (int i = 0; i < 8; i++) w[i] = a[i];
  int m = 0;
#pragma unroll
  for (int i = 0; i < 7; i++) {
#pragma unroll
    for (int j = i + 1; j < 8; j += 2) {
      const int h0 = (int)__umulhi(w[i], w[j]);
      const int h1 = (j + 1 < 8) ? (int)__umulhi(w[i], w[j + 1]) : 0;
      m = __vimax3_s32(m, h0, h1);
    }
  }
  return (uint32_t)m == 0x7fffffffu;
}

// First-level filter.  Eight of the 28 cross products (ECB200_SQR_EXACT_PAIRS, written by gen_fp256.py: the two
// chains of fp_sqr_t9 whose products all land on untouched accumulator pairs) leave their exact high words
// in registers for free: fp_sqr_t9 folds them into qx (signed max, INT_MAX <=> hit).  The other 20 get
// a necessary condition in fp32, one FFMA each: a hit needs a_i, a_j >= 2^31 and
// a_i*a_j in [2^63 - 2^32, 2^63).  f(a) = as_float(0x3f000000 + (a >> 8)) is one LEA.HI; for a >= 2^31
// it equals floor(a / 256) * 2^-23, i.e. a / 2^31 truncated to 24 bits, in [1, 2).  Then
//     2 - 2^-30 - 3 * 2^-23  <  f_i * f_j  <=  a_i * a_j / 2^62  <  2
// (x + y < 3 for x, y in [1, 2) with x*y < 2), so d = fma(-f_i, f_j, 2) -- one rounding of the exact
// value -- lies in [0, 2^-21]: as an unsigned integer, bits(d) <= 0x35000000; negative d (product
// above 2) has bits >= 0x80000000 and limbs below 2^31 can only add false positives.  An unsigned
// 3-input min keeps it to one FFMA + half a VIMNMX3 per pair; `m` chains through a group of squarings.
// False positives: ~2e-6 per square (the 16-bit integer form this replaces: 8.5e-4, i.e. a cold-path
// excursion in one ladder step out of six per warp).
__host__ __device__ constexpr bool fp_sqr_pair_is_exact(int i, int j) {
  constexpr int ex[8][2] = ECB200_SQR_EXACT_PAIRS;
  for (int k = 0; k < 8; k++)
    if (ex[k][0] == i && ex[k][1] == j) return true;
  return false;
}
#define ECB200_QF_THRESHOLD 0x35000000u
__device__ __forceinline__ uint32_t fp_sqr_quirk_filter(const fe& a, uint32_t m = 0xffffffffu) {
  float f[8];
#pragma unroll
  for (int i = 0; i < 8; i++) f[i] = __uint_as_float((a.v[i] >> 8) + 0x3f000000u);
  uint32_t pend = 0xffffffffu;
  bool have = false;
#pragma unroll
  for (int i = 0; i < 7; i++) {
#pragma unroll
    for (int j = i + 1; j < 8; j++) {
      if (fp_sqr_pair_is_exact(i, j)) continue;  // in qx
      const uint32_t y = __float_as_uint(__fmaf_rn(-f[i], f[j], 2.0f));
      if (have) { m = __vimin3_u32(m, pend, y); have = false; }
      else { pend = y; have = true; }
    }
  }
  if (have) m = min(m, pend);
  return m;
}
// hit test of a group: some approximate pair below the threshold, or some exact pair at INT_MAX
__device__ __forceinline__ bool fp_quirk_maybe(uint32_t fm, uint32_t qx) { return fm <= ECB200_QF_THRESHOLD || qx == 0x7fffffffu; }
#define ECB200_QX_INIT 0x80000000u  /* INT_MIN: neutral element of the signed max */
// all 28 pairs through the IMAD condition (callers that do not run fp_sqr_t9: the generic-prime path)
__device__ __forceinline__ uint32_t fp_sqr_quirk_filter_all(const fe& a) {
  uint32_t h[8];
#pragma unroll
  for (int i = 0; i < 8; i++) h[i] = a.v[i] >> 16;
  uint32_t m = 0xffffffffu;
#pragma unroll
  for (int i = 0; i < 7; i++) {
#pragma unroll
    for (int j = i + 1; j < 8; j += 2) {
      const uint32_t y0 = h[i] * h[j] + 0x80020000u;
      const uint32_t y1 = (j + 1 < 8) ? h[i] * h[j + 1] + 0x80020000u : 0xffffffffu;
      m = __vimin3_u32(m, y0, y1);
    }
  }
  return m;
}
static_assert(sizeof((const int[][2])ECB200_SQR_EXACT_PAIRS) == 8 * 2 * sizeof(int), "fp_sqr_pair_is_exact reads exactly 8 exact pairs");

template <bool QUIRK, class M>
__device__ __forceinline__ fe fp_sqr_core(const fe& a, M& mode, uint32_t& qx) {
  fe t;
  uint32_t t8;
  fp_sqr_t9(t.v[0], t.v[1], t.v[2], t.v[3], t.v[4], t.v[5], t.v[6], t.v[7], t8, qx,
            a.v[0], a.v[1], a.v[2], a.v[3], a.v[4], a.v[5], a.v[6], a.v[7]);
  return fp_reduce_once(t, t8, mode);
}

template <bool QUIRK>
__device__ __forceinline__ fe fp_sqr(const fe& a, Exact& mode) {
  uint32_t qx = ECB200_QX_INIT;
  fe r = fp_sqr_core<QUIRK>(a, mode, qx);
  if (QUIRK) {
    if (__builtin_expect(fp_quirk_maybe(fp_sqr_quirk_filter(a), qx), 0)) {
      uint32_t in[8], out[8];
#pragma unroll
      for (int i = 0; i < 8; i++) in[i] = a.v[i];
      if (fp_sqr_quirk_filter_exact(in)) {
        fp_sqr_quirk_slow(out, in);
#pragma unroll
        for (int i = 0; i < 8; i++) r.v[i] = out[i];
      }
    }
  }
  return r;
}

// Cold path shared by the Lazy-mode squarings: for each of the n operands (8 words each) that
// passes the exact test, overwrite its result with the reference's defective square.
static __device__ __noinline__ uint32_t fp_quirk_fix(uint32_t* res, const uint32_t* ops, int n) {
  uint32_t fixed = 0;
  for (int k = 0; k < n; k++) {
    if (fp_sqr_quirk_filter_exact(ops + 8 * k)) {
      fp_sqr_quirk_slow(res + 8 * k, ops + 8 * k);
      fixed |= 1u << k;
    }
  }
  return fixed;
}

template <bool QUIRK>
__device__ __forceinline__ fe fp_sqr(const fe& a, Lazy& mode) {
  uint32_t qx = ECB200_QX_INIT;
  fe r = fp_sqr_core<QUIRK>(a, mode, qx);
  if (QUIRK) {
    if (__builtin_expect(fp_quirk_maybe(fp_sqr_quirk_filter(a), qx), 0)) {
      uint32_t in[8], out[8];
#pragma unroll
      for (int i = 0; i < 8; i++) in[i] = a.v[i];
      if (fp_quirk_fix(out, in, 1)) {
#pragma unroll
        for (int i = 0; i < 8; i++) r.v[i] = out[i];
      }
    }
  }
  return r;
}
template <bool QUIRK = true>
__device__ __forceinline__ fe fp_sqr(const fe& a) { Exact e; return fp_sqr<QUIRK>(a, e); }

// Grouped form for the point formulas: fp_sqr_acc squares and only folds the first-level filter of
// its operand into the group's accumulators (QuirkAcc: unsigned min of the fp32 test, signed max of
// the exact high words); fp_quirk_check then resolves a whole group of squarings with ONE branch,
// repairing the results in place -- it must follow the group's squarings before any of their
// results is used.  Exact mode resolves each squaring in place (accumulators unused), so both modes
// share the formulas.
struct QuirkAcc {
  uint32_t fm = 0xffffffffu;
  uint32_t qx = ECB200_QX_INIT;
};
template <bool QUIRK>
__device__ __forceinline__ fe fp_sqr_acc(const fe& a, Exact& mode, QuirkAcc&) { return fp_sqr<QUIRK>(a, mode); }
template <bool QUIRK>
__device__ __forceinline__ fe fp_sqr_acc(const fe& a, Lazy& mode, QuirkAcc& q) {
  const fe r = fp_sqr_core<QUIRK>(a, mode, q.qx);
  if (QUIRK) q.fm = fp_sqr_quirk_filter(a, q.fm);
  return r;
}
template <bool QUIRK>
__device__ __forceinline__ void fp_quirk_check(Exact&, const QuirkAcc&, const fe&, fe&, const fe&, fe&) {}
template <bool QUIRK>
__device__ __forceinline__ void fp_quirk_check(Exact&, const QuirkAcc&, const fe&, fe&, const fe&, fe&, const fe&, fe&) {}
template <bool QUIRK>
__device__ __forceinline__ void fp_quirk_check(Lazy&, const QuirkAcc& q, const fe& a, fe& ra, const fe& b, fe& rb) {
  if (QUIRK) {
    if (__builtin_expect(fp_quirk_maybe(q.fm, q.qx), 0)) {
      // volatile: word-by-word local stores.  128-bit ones would force the operands and results of every squaring
      // into aligned register quads on the HOT path (consecutive registers = alternating banks, sass_recolor.py)
      volatile uint32_t in[16], out[16];
#pragma unroll
      for (int i = 0; i < 8; i++) { in[i] = a.v[i]; in[8 + i] = b.v[i]; out[i] = ra.v[i]; out[8 + i] = rb.v[i]; }
      if (fp_quirk_fix(const_cast<uint32_t*>(out), const_cast<const uint32_t*>(in), 2)) {
#pragma unroll
        for (int i = 0; i < 8; i++) { ra.v[i] = out[i]; rb.v[i] = out[8 + i]; }
      }
    }
  }
}
template <bool QUIRK>
__device__ __forceinline__ void fp_quirk_check(Lazy&, const QuirkAcc& q, const fe& a, fe& ra, const fe& b, fe& rb, const fe& c, fe& rc) {
  if (QUIRK) {
    if (__builtin_expect(fp_quirk_maybe(q.fm, q.qx), 0)) {
      volatile uint32_t in[24], out[24];
#pragma unroll
      for (int i = 0; i < 8; i++) {
        in[i] = a.v[i]; in[8 + i] = b.v[i]; in[16 + i] = c.v[i];
        out[i] = ra.v[i]; out[8 + i] = rb.v[i]; out[16 + i] = rc.v[i];
      }
      if (fp_quirk_fix(const_cast<uint32_t*>(out), const_cast<const uint32_t*>(in), 3)) {
#pragma unroll
        for (int i = 0; i < 8; i++) { ra.v[i] = out[i]; rb.v[i] = out[8 + i]; rc.v[i] = out[16 + i]; }
      }
    }
  }
}

// a^e through the reference's LSB-first square-and-multiply (mgry_ops.h:44-86): bit b of e multiplies the running
// product by a^(2^b), and the base is squared nbits-1 times -- the same sequence of squarings (hence the same
// squaring-defect lanes) as the reference.  e is a compile-time constant of the callers (p-2, (p+1)/4): the branch
// on its bits is uniform.  MD = Lazy keeps the 2^-32 corner cases out of the loop (the caller re-runs a flagged
// lane in Exact mode), which is what lets the loop body stay branch-free apart from the defect filter.
template <bool QUIRK, class MD>
__device__ __forceinline__ fe fp_pow_lsb(const fe& a, const uint32_t (&e)[8], int nbits, MD& md) {
  fe res = fe_R(), base = a;
#pragma unroll 1
  for (int b = 0; b < nbits; b++) {
    if ((e[b >> 5] >> (b & 31)) & 1u) res = fp_mul(res, base, md);
    if (b < nbits - 1) base = fp_sqr<QUIRK>(base, md);
  }
  return res;
}
#define ECB200_PM2_WORDS {0xfffffffdu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u, 1u, 0xffffffffu}            /* p - 2 */
#define ECB200_SQRT_EXP_WORDS {0u, 0u, 0x40000000u, 0u, 0u, 0x40000000u, 0xc0000000u, 0x3fffffffu}    /* (p+1)/4 = 2^254 - 2^222 + 2^190 + 2^94 */
// GFp::inverse = a^(p-2)   gfp.h:42-44
template <bool QUIRK, class MD>
__device__ __forceinline__ fe fp_inv(const fe& a, MD& md) {
  const uint32_t e[8] = ECB200_PM2_WORDS;
  return fp_pow_lsb<QUIRK>(a, e, 256, md);
}

// gfp.h:60-64: opposite(a) = (p-1)R - (a - R)
__device__ __forceinline__ fe fp_neg(const fe& a) { return fp_sub(fe_PM1R(), fp_sub(a, fe_R())); }

// mgry.h:47-50 / :52-55
template <class MD>
__device__ __forceinline__ fe fp_from_classical(const fe& a, MD& md) { return fp_mul(a, fe_RR(), md); }
template <class MD>
__device__ __forceinline__ fe fp_to_classical(const fe& a, MD& md) {
  fe one = fe_zero();
  one.v[0] = 1;
  return fp_mul(a, one, md);  // (a * 1 + m p) / R == mgry_reduce(pad(a))
}
__device__ __forceinline__ fe fp_from_classical(const fe& a) { Exact e; return fp_from_classical(a, e); }
__device__ __forceinline__ fe fp_to_classical(const fe& a) { Exact e; return fp_to_classical(a, e); }

__device__ __forceinline__ void fe_cswap(uint32_t mask, fe& a, fe& b) {
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const uint32_t x = a.v[i], y = b.v[i];
    a.v[i] = mask ? y : x;
    b.v[i] = mask ? x : y;
  }
}
__device__ __forceinline__ bool fe_eq(const fe& a, const fe& b) {
  uint32_t d = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) d |= a.v[i] ^ b.v[i];
  return d == 0;
}

}  // namespace ecb200
