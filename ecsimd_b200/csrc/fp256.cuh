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

// s (with carry-out c): return s if (s - p borrows and c == 0) else s - p
__device__ __forceinline__ fe fp_reduce_once(const fe& s, uint32_t c) {
  fe d;
  uint32_t bw;
  asm("sub.cc.u32 %0, %9, 0xffffffff; subc.cc.u32 %1, %10, 0xffffffff; subc.cc.u32 %2, %11, 0xffffffff; subc.cc.u32 %3, %12, 0; "
      "subc.cc.u32 %4, %13, 0; subc.cc.u32 %5, %14, 0; subc.cc.u32 %6, %15, 1; subc.cc.u32 %7, %16, 0xffffffff; "
      "subc.u32 %8, 0, 0;"
      : "=r"(d.v[0]), "=r"(d.v[1]), "=r"(d.v[2]), "=r"(d.v[3]), "=r"(d.v[4]), "=r"(d.v[5]), "=r"(d.v[6]), "=r"(d.v[7]), "=r"(bw)
      : "r"(s.v[0]), "r"(s.v[1]), "r"(s.v[2]), "r"(s.v[3]), "r"(s.v[4]), "r"(s.v[5]), "r"(s.v[6]), "r"(s.v[7]));
  const bool keep = (bw != 0u) && (c == 0u);
  fe r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = keep ? s.v[i] : d.v[i];
  return r;
}

__device__ __forceinline__ fe fp_add(const fe& a, const fe& b) {
  fe s;
  uint32_t c;
  asm("add.cc.u32 %0, %9, %17; addc.cc.u32 %1, %10, %18; addc.cc.u32 %2, %11, %19; addc.cc.u32 %3, %12, %20; "
      "addc.cc.u32 %4, %13, %21; addc.cc.u32 %5, %14, %22; addc.cc.u32 %6, %15, %23; addc.cc.u32 %7, %16, %24; "
      "addc.u32 %8, 0, 0;"
      : "=r"(s.v[0]), "=r"(s.v[1]), "=r"(s.v[2]), "=r"(s.v[3]), "=r"(s.v[4]), "=r"(s.v[5]), "=r"(s.v[6]), "=r"(s.v[7]), "=r"(c)
      : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
        "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
  return fp_reduce_once(s, c);
}

__device__ __forceinline__ fe fp_shl1(const fe& a) {
  fe s;
  const uint32_t c = a.v[7] >> 31;
#pragma unroll
  for (int i = 7; i > 0; i--) s.v[i] = __funnelshift_l(a.v[i - 1], a.v[i], 1);
  s.v[0] = a.v[0] << 1;
  return fp_reduce_once(s, c);
}
template <int COUNT>
__device__ __forceinline__ fe fp_shl(const fe& a) {
  fe r = a;
#pragma unroll
  for (int i = 0; i < COUNT; i++) r = fp_shl1(r);
  return r;
}

__device__ __forceinline__ fe fp_mul(const fe& a, const fe& b) {
  fe r;
  fp_mul_words(r.v[0], r.v[1], r.v[2], r.v[3], r.v[4], r.v[5], r.v[6], r.v[7],
               a.v[0], a.v[1], a.v[2], a.v[3], a.v[4], a.v[5], a.v[6], a.v[7],
               b.v[0], b.v[1], b.v[2], b.v[3], b.v[4], b.v[5], b.v[6], b.v[7]);
  return r;
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
// is INT_MAX.  So: fast path = true square (the multiplier above), plus a
// 3-input signed max over the 28 cross-product high words; only if that max is
// INT_MAX (probability ~ 6.5e-9 per lane) the lane re-runs the reference's loop
// literally (fp_sqr_quirk_slow).
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

__device__ __forceinline__ uint32_t fp_sqr_quirk_filter(const fe& a) {
  int m = 0;
#pragma unroll
  for (int i = 0; i < 7; i++) {
#pragma unroll
    for (int j = i + 1; j < 8; j += 2) {
      const int h0 = (int)__umulhi(a.v[i], a.v[j]);
      const int h1 = (j + 1 < 8) ? (int)__umulhi(a.v[i], a.v[j + 1]) : 0;
      m = __vimax3_s32(m, h0, h1);
    }
  }
  return (uint32_t)m;
}

template <bool QUIRK = true>
__device__ __forceinline__ fe fp_sqr(const fe& a) {
  fe r = fp_mul(a, a);
  if (QUIRK) {
    if (fp_sqr_quirk_filter(a) == 0x7fffffffu) {
      uint32_t in[8], out[8];
#pragma unroll
      for (int i = 0; i < 8; i++) in[i] = a.v[i];
      fp_sqr_quirk_slow(out, in);
#pragma unroll
      for (int i = 0; i < 8; i++) r.v[i] = out[i];
    }
  }
  return r;
}

// gfp.h:60-64: opposite(a) = (p-1)R - (a - R)
__device__ __forceinline__ fe fp_neg(const fe& a) { return fp_sub(fe_PM1R(), fp_sub(a, fe_R())); }

// mgry.h:47-50 / :52-55
__device__ __forceinline__ fe fp_from_classical(const fe& a) { return fp_mul(a, fe_RR()); }
__device__ __forceinline__ fe fp_to_classical(const fe& a) {
  fe one = fe_zero();
  one.v[0] = 1;
  return fp_mul(a, one);  // (a * 1 + m p) / R == mgry_reduce(pad(a))
}

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
