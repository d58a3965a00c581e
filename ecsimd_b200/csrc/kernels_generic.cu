// kernels_generic.cu -- run-time-modulus field kernels and their C-ABI entry points (SURVEY 8f-4:
// lets the reference's own secp256k1 field tests, tests/mgry.cpp and tests/ops.cpp Ops256.*, run
// against the engine).
#include "fpgen.cuh"
#include "host_common.cuh"
#include "layout.cuh"

namespace ecb200 {

enum GenOp : int { G_ADD, G_SUB, G_SHL1, G_MUL, G_SQR, G_FROMC, G_TOC, G_POW, G_OPP };

template <int L, int OP, bool QUIRK>
__global__ void __launch_bounds__(128) k_gen(void* __restrict__ out, const void* __restrict__ a, const void* __restrict__ b,
                                             size_t n, GenPrime P, fe e) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const fe x = Layout<L>::load(a, n, i, 1, 0);
  fe r;
  if (OP == G_ADD) r = gen_add(x, Layout<L>::load(b, n, i, 1, 0), P);
  else if (OP == G_SUB) r = gen_sub(x, Layout<L>::load(b, n, i, 1, 0), P);
  else if (OP == G_SHL1) r = gen_shl1(x, P);
  else if (OP == G_MUL) r = gen_mul(x, Layout<L>::load(b, n, i, 1, 0), P);
  else if (OP == G_SQR) r = gen_sqr<QUIRK>(x, P);
  else if (OP == G_FROMC) r = gen_mul(x, fe_const(P.rr), P);                       // mgry.h:47-50
  else if (OP == G_TOC) { fe one = fe_zero(); one.v[0] = 1; r = gen_mul(x, one, P); }  // mgry.h:52-55
  else if (OP == G_POW) r = gen_pow<QUIRK>(x, e.v, P);
  else {  // gfp.h:60-64: opposite(a) = (p-1)R - (a - R); (p-1)R mod p = p - R mod p
    fe pm1r, pp = fe_const(P.p), r1 = fe_const(P.r1);
    uint32_t bw = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const unsigned long long y = (unsigned long long)pp.v[k] - r1.v[k] - bw;
      pm1r.v[k] = (uint32_t)y;
      bw = (uint32_t)(y >> 63);
    }
    r = gen_sub(pm1r, gen_sub(x, r1, P), P);
  }
  Layout<L>::store(out, n, i, 1, 0, r);
}

// 256x256 -> 512 (mul.h:150-158) and the reference's square() (mul.h:214-221); lane layout, 16 words out
template <bool SQUARE>
__global__ void __launch_bounds__(128) k_wide512(uint32_t* __restrict__ out16, const uint32_t* __restrict__ a, const uint32_t* __restrict__ b, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t x[8], t[16];
  for (int k = 0; k < 8; k++) x[k] = a[i * 8 + k];
  if (SQUARE) square512_quirk(t, x);
  else {
    uint32_t y[8];
    for (int k = 0; k < 8; k++) y[k] = b[i * 8 + k];
    fp_mul512_words(t[0], t[1], t[2], t[3], t[4], t[5], t[6], t[7], t[8], t[9], t[10], t[11], t[12], t[13], t[14], t[15],
                    x[0], x[1], x[2], x[3], x[4], x[5], x[6], x[7], y[0], y[1], y[2], y[3], y[4], y[5], y[6], y[7]);
  }
  for (int k = 0; k < 16; k++) out16[i * 16 + k] = t[k];
}

// ---- host: derive the Montgomery constants of p -----------------------------------------------------------
static bool ge8(const uint32_t* a, const uint32_t* p) {
  for (int i = 7; i >= 0; i--)
    if (a[i] != p[i]) return a[i] > p[i];
  return true;
}
static void dbl_mod(uint32_t* x, const uint32_t* p) {  // x = 2x mod p, x < p
  uint32_t c = x[7] >> 31;
  for (int i = 7; i > 0; i--) x[i] = (x[i] << 1) | (x[i - 1] >> 31);
  x[0] <<= 1;
  if (c || ge8(x, p)) {
    unsigned long long bw = 0;
    for (int i = 0; i < 8; i++) {
      const unsigned long long d = (unsigned long long)x[i] - p[i] - bw;
      x[i] = (uint32_t)d;
      bw = (d >> 63) & 1;
    }
  }
}
static int make_prime(GenPrime* P, const uint32_t* p8) {
  if (!p8 || !(p8[0] & 1u) || !(p8[7] >> 31)) {
    set_error("generic modulus must be odd with bit 255 set (a 256-bit modulus, R = 2^256)");
    return ECB200_ERR_ARG;
  }
  std::memcpy(P->p, p8, 32);
  uint32_t inv = p8[0];  // Newton: inv = p^-1 mod 2^32
  for (int i = 0; i < 5; i++) inv *= 2u - p8[0] * inv;
  P->mprime = 0u - inv;
  uint32_t x[8] = {1, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 256; i++) dbl_mod(x, p8);
  std::memcpy(P->r1, x, 32);
  for (int i = 0; i < 256; i++) dbl_mod(x, p8);
  std::memcpy(P->rr, x, 32);
  return ECB200_OK;
}

template <int OP, bool Q>
static int launch_gen_l(int L, void* out, const void* a, const void* b, size_t n, const GenPrime& P, const fe& e, cudaStream_t s) {
  const unsigned blocks = (unsigned)((n + 127) / 128);
  if (L == L_LANE) k_gen<L_LANE, OP, Q><<<blocks, 128, 0, s>>>(out, a, b, n, P, e);
  else if (L == L_PACK4) k_gen<L_PACK4, OP, Q><<<blocks, 128, 0, s>>>(out, a, b, n, P, e);
  else k_gen<L_SOA, OP, Q><<<blocks, 128, 0, s>>>(out, a, b, n, P, e);
  ECB_LAUNCH_CHECK();
  return ECB200_OK;
}

template <int OP>
static int gen_call(void* out, const void* a, const void* b, const uint32_t* e8, const uint32_t* p8, size_t n, uint32_t flags, void* stream) {
  int rc = check_common(n, flags);
  if (rc) return rc;
  GenPrime P;
  if ((rc = make_prime(&P, p8))) return rc;
  if (n == 0) return ECB200_OK;
  const bool binary = (OP == G_ADD || OP == G_SUB || OP == G_MUL);
  if (!out || !a || (binary && !b) || (OP == G_POW && !e8)) { set_error("null pointer argument"); return ECB200_ERR_ARG; }
  fe e = {};
  if (e8) std::memcpy(e.v, e8, 32);
  cudaStream_t s = (cudaStream_t)stream;
  const int L = layout_of(flags);
  const size_t bytes = operand_bytes(n, 1);
  Scratch sc(s);
  const void *da = a, *db = b;
  void* dout = out;
  if (!on_device(flags)) {
    void *x, *y = nullptr, *z;
    if ((rc = sc.alloc(&x, bytes)) || (rc = sc.alloc(&z, bytes))) return rc;
    ECB_CUDA(cudaMemcpyAsync(x, a, bytes, cudaMemcpyHostToDevice, s));
    if (binary) {
      if ((rc = sc.alloc(&y, bytes))) return rc;
      ECB_CUDA(cudaMemcpyAsync(y, b, bytes, cudaMemcpyHostToDevice, s));
    }
    da = x; db = y; dout = z;
  }
  const bool q = quirk_on(flags);
  rc = (OP == G_SQR || OP == G_POW) && !q ? launch_gen_l<OP, false>(L, dout, da, db, n, P, e, s) : launch_gen_l<OP, true>(L, dout, da, db, n, P, e, s);
  if (rc) return rc;
  if (!on_device(flags)) {
    ECB_CUDA(cudaMemcpyAsync(out, dout, bytes, cudaMemcpyDeviceToHost, s));
    ECB_CUDA(cudaStreamSynchronize(s));
  }
  return ECB200_OK;
}

template <bool SQUARE>
static int wide_call(void* out16, const void* a, const void* b, size_t n, uint32_t flags, void* stream) {
  if (layout_of(flags) != L_LANE) { set_error("512-bit products use ECB200_LAYOUT_LANE"); return ECB200_ERR_ARG; }
  if (n == 0) return ECB200_OK;
  if (!out16 || !a || (!SQUARE && !b)) { set_error("null pointer argument"); return ECB200_ERR_ARG; }
  cudaStream_t s = (cudaStream_t)stream;
  Scratch sc(s);
  int rc;
  const void *da = a, *db = b;
  void* dout = out16;
  if (!on_device(flags)) {
    void *x, *y = nullptr, *z;
    if ((rc = sc.alloc(&x, n * 32)) || (rc = sc.alloc(&z, n * 64))) return rc;
    ECB_CUDA(cudaMemcpyAsync(x, a, n * 32, cudaMemcpyHostToDevice, s));
    if (!SQUARE) {
      if ((rc = sc.alloc(&y, n * 32))) return rc;
      ECB_CUDA(cudaMemcpyAsync(y, b, n * 32, cudaMemcpyHostToDevice, s));
    }
    da = x; db = y; dout = z;
  }
  k_wide512<SQUARE><<<(unsigned)((n + 127) / 128), 128, 0, s>>>((uint32_t*)dout, (const uint32_t*)da, (const uint32_t*)db, n);
  ECB_LAUNCH_CHECK();
  if (!on_device(flags)) {
    ECB_CUDA(cudaMemcpyAsync(out16, dout, n * 64, cudaMemcpyDeviceToHost, s));
    ECB_CUDA(cudaStreamSynchronize(s));
  }
  return ECB200_OK;
}

}  // namespace ecb200

using namespace ecb200;

extern "C" {
int ecb200_gen_mod_add(void* out, const void* a, const void* b, const uint32_t* p8, size_t n, uint32_t flags, void* stream) { return gen_call<G_ADD>(out, a, b, nullptr, p8, n, flags, stream); }
int ecb200_gen_mod_sub(void* out, const void* a, const void* b, const uint32_t* p8, size_t n, uint32_t flags, void* stream) { return gen_call<G_SUB>(out, a, b, nullptr, p8, n, flags, stream); }
int ecb200_gen_mod_shift_left_one(void* out, const void* a, const uint32_t* p8, size_t n, uint32_t flags, void* stream) { return gen_call<G_SHL1>(out, a, nullptr, nullptr, p8, n, flags, stream); }
int ecb200_gen_mgry_mul(void* out, const void* a, const void* b, const uint32_t* p8, size_t n, uint32_t flags, void* stream) { return gen_call<G_MUL>(out, a, b, nullptr, p8, n, flags, stream); }
int ecb200_gen_mgry_sqr(void* out, const void* a, const uint32_t* p8, size_t n, uint32_t flags, void* stream) { return gen_call<G_SQR>(out, a, nullptr, nullptr, p8, n, flags, stream); }
int ecb200_gen_from_classical(void* out, const void* a, const uint32_t* p8, size_t n, uint32_t flags, void* stream) { return gen_call<G_FROMC>(out, a, nullptr, nullptr, p8, n, flags, stream); }
int ecb200_gen_to_classical(void* out, const void* a, const uint32_t* p8, size_t n, uint32_t flags, void* stream) { return gen_call<G_TOC>(out, a, nullptr, nullptr, p8, n, flags, stream); }
int ecb200_gen_mgry_pow(void* out, const void* a, const uint32_t* e8, const uint32_t* p8, size_t n, uint32_t flags, void* stream) { return gen_call<G_POW>(out, a, nullptr, e8, p8, n, flags, stream); }
int ecb200_gen_opposite(void* out, const void* a, const uint32_t* p8, size_t n, uint32_t flags, void* stream) { return gen_call<G_OPP>(out, a, nullptr, nullptr, p8, n, flags, stream); }
int ecb200_mul512(void* out16, const void* a, const void* b, size_t n, uint32_t flags, void* stream) { return wide_call<false>(out16, a, b, n, flags, stream); }
int ecb200_square512(void* out16, const void* a, size_t n, uint32_t flags, void* stream) { return wide_call<true>(out16, a, nullptr, n, flags, stream); }
}
