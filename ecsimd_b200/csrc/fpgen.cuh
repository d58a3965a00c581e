// fpgen.cuh -- Montgomery field arithmetic for an arbitrary odd 256-bit modulus given at run time.
//
// The reference's field layer is a template over the prime (include/ecsimd/mgry_ops.h,
// modular.h, mgry_mul.h); its own field tests run it on the secp256k1 prime
// (tests/mgry.cpp:25-27, tests/ops.cpp:221-252).  This is the same layer with the modulus as data:
// correct (bit-exact with the reference for any 256-bit input, including the squaring defect)
// rather than tuned -- the tuned path is the P-256 specialisation in fp256.cuh.
#pragma once
#include "fp256.cuh"

namespace ecb200 {

struct GenPrime {
  uint32_t p[8];    // modulus, least-significant word first
  uint32_t r1[8];   // R mod p            (mgry_csts.h:20)
  uint32_t rr[8];   // R^2 mod p          (mgry_csts.h:21)
  uint32_t mprime;  // -p^-1 mod 2^32     (mgry_mul.h:33-40)
};

// sub_if_above (sub.h:46-69) for modulus P: keep s iff (s - p borrows) and no carry came in
__device__ __forceinline__ fe gen_reduce_once(const fe& s, uint32_t c, const GenPrime& P) {
  fe d;
  uint32_t bw = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const unsigned long long x = (unsigned long long)s.v[i] - P.p[i] - bw;
    d.v[i] = (uint32_t)x;
    bw = (uint32_t)(x >> 63);
  }
  const bool keep = bw && !c;
  fe r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = keep ? s.v[i] : d.v[i];
  return r;
}
// modular.h:10-15
__device__ __forceinline__ fe gen_add(const fe& a, const fe& b, const GenPrime& P) {
  fe s;
  uint32_t c = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const unsigned long long x = (unsigned long long)a.v[i] + b.v[i] + c;
    s.v[i] = (uint32_t)x;
    c = (uint32_t)(x >> 32);
  }
  return gen_reduce_once(s, c, P);
}
// modular.h:24-41
__device__ __forceinline__ fe gen_sub(const fe& a, const fe& b, const GenPrime& P) {
  fe d;
  uint32_t bw = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const unsigned long long x = (unsigned long long)a.v[i] - b.v[i] - bw;
    d.v[i] = (uint32_t)x;
    bw = (uint32_t)(x >> 63);
  }
  uint32_t c = 0;
  fe r;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const unsigned long long x = (unsigned long long)d.v[i] + (bw ? P.p[i] : 0u) + c;
    r.v[i] = (uint32_t)x;
    c = (uint32_t)(x >> 32);
  }
  return r;
}
// modular.h:17-22
__device__ __forceinline__ fe gen_shl1(const fe& a, const GenPrime& P) {
  fe s;
  const uint32_t c = a.v[7] >> 31;
#pragma unroll
  for (int i = 7; i > 0; i--) s.v[i] = __funnelshift_l(a.v[i - 1], a.v[i], 1);
  s.v[0] = a.v[0] << 1;
  return gen_reduce_once(s, c, P);
}

// mgry_reduce (mgry_mul.h:84-121) of a 16-word T: t = (T + m p) / 2^256 word by word, then minus p iff t >= p
__device__ __forceinline__ fe gen_redc(uint32_t (&t)[16], const GenPrime& P) {
  uint32_t top = 0;  // word 16
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const uint32_t m = t[i] * P.mprime;
    unsigned long long carry = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const unsigned long long x = (unsigned long long)m * P.p[k] + t[i + k] + carry;
      t[i + k] = (uint32_t)x;
      carry = x >> 32;
    }
#pragma unroll
    for (int k = i + 8; k < 16; k++) {
      const unsigned long long x = (unsigned long long)t[k] + carry;
      t[k] = (uint32_t)x;
      carry = x >> 32;
    }
    top += (uint32_t)carry;
  }
  fe s;
#pragma unroll
  for (int i = 0; i < 8; i++) s.v[i] = t[8 + i];
  return gen_reduce_once(s, top, P);
}

__device__ __forceinline__ fe gen_mul(const fe& a, const fe& b, const GenPrime& P) {
  uint32_t t[16];
  fp_mul512_words(t[0], t[1], t[2], t[3], t[4], t[5], t[6], t[7], t[8], t[9], t[10], t[11], t[12], t[13], t[14], t[15],
                  a.v[0], a.v[1], a.v[2], a.v[3], a.v[4], a.v[5], a.v[6], a.v[7],
                  b.v[0], b.v[1], b.v[2], b.v[3], b.v[4], b.v[5], b.v[6], b.v[7]);
  return gen_redc(t, P);
}

// the reference's square() with its lost-carry defect (mul.h:160-221), 16 words out
static __device__ __noinline__ void square512_quirk(uint32_t* r16, const uint32_t* a) {
  unsigned long long ret[17];
  for (int k = 0; k < 17; k++) ret[k] = 0;
  for (int i = 0; i < 8; i++) {
    unsigned long long t = (unsigned long long)a[i] * a[i] + ret[2 * i];
    ret[2 * i] = t & 0xffffffffull;
    unsigned long long p0 = t >> 32, p1 = 0;
    for (int j = i + 1; j < 8; j++) {
      const unsigned long long pr = (unsigned long long)a[i] * a[j];
      const unsigned long long carry = pr >> 63;
      t = (pr << 1) + ret[i + j] + p0;  // may wrap: the defect being reproduced
      ret[i + j] = t & 0xffffffffull;
      p0 = p1 + (t >> 32);
      p1 = carry;
    }
    ret[i + 8] += p0;
    if (i + 9 < 16) ret[i + 9] = p1;
  }
  for (int k = 0; k < 16; k++) r16[k] = (uint32_t)ret[k];
}

template <bool QUIRK>
__device__ __forceinline__ fe gen_sqr(const fe& a, const GenPrime& P) {
  uint32_t t[16];
  fp_mul512_words(t[0], t[1], t[2], t[3], t[4], t[5], t[6], t[7], t[8], t[9], t[10], t[11], t[12], t[13], t[14], t[15],
                  a.v[0], a.v[1], a.v[2], a.v[3], a.v[4], a.v[5], a.v[6], a.v[7],
                  a.v[0], a.v[1], a.v[2], a.v[3], a.v[4], a.v[5], a.v[6], a.v[7]);
  if (QUIRK) {
    if (__builtin_expect(fp_sqr_quirk_filter_all(a) < 0x20000u, 0)) {
      uint32_t in[8];
#pragma unroll
      for (int i = 0; i < 8; i++) in[i] = a.v[i];
      if (fp_sqr_quirk_filter_exact(in)) {
        uint32_t q[16];
        square512_quirk(q, in);
#pragma unroll
        for (int i = 0; i < 16; i++) t[i] = q[i];
      }
    }
  }
  return gen_redc(t, P);
}

// mgry_pow (mgry_ops.h:44-86): LSB-first square-and-multiply, no squaring after the top set bit
template <bool QUIRK>
__device__ __forceinline__ fe gen_pow(const fe& a, const uint32_t (&e)[8], const GenPrime& P) {
  int top = -1;
  for (int b = 255; b >= 0; b--)
    if ((e[b >> 5] >> (b & 31)) & 1u) { top = b; break; }
  fe res = fe_const(P.r1), base = a;
#pragma unroll 1
  for (int b = 0; b <= top; b++) {
    if ((e[b >> 5] >> (b & 31)) & 1u) res = gen_mul(res, base, P);
    if (b < top) base = gen_sqr<QUIRK>(base, P);
  }
  return res;
}

}  // namespace ecb200
