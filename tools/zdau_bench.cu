// tools/zdau_bench.cu -- development harness: times the ZDAU ladder step (the 100 % hot loop of
// scalar_mult, curve_group.h:198-212) in isolation, in several code-shape variants.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/zdau_bench tools/zdau_bench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../ecsimd_b200/csrc/point.cuh"
using namespace ecb200;

#ifndef VARIANT_CALLS
#define VARIANT_CALLS 1
#endif

__device__ __noinline__ fe mul_call(fe a, fe b) { return fp_mul(a, b); }
__device__ __noinline__ fe sqr_call(fe a) { return fp_sqr<true>(a); }

template <bool CALLS>
__device__ __forceinline__ fe MUL(const fe& a, const fe& b) { return CALLS ? mul_call(a, b) : fp_mul(a, b); }
#ifndef BENCH_QUIRK
#define BENCH_QUIRK true
#endif
template <bool CALLS>
__device__ __forceinline__ fe SQR(const fe& a) { return CALLS ? sqr_call(a) : fp_sqr<BENCH_QUIRK>(a); }

template <bool CALLS>
__device__ __forceinline__ void zdau_v(fe& X1, fe& Y1, fe& X2, fe& Y2, fe& Z) {
  const fe dx = fp_sub(X1, X2);
  const fe dy = fp_sub(Y1, Y2);
  const fe Cp = SQR<CALLS>(dx);
  const fe W1p = MUL<CALLS>(X1, Cp);
  const fe W2p = MUL<CALLS>(X2, Cp);
  const fe Dp = SQR<CALLS>(dy);
  const fe A1p = MUL<CALLS>(Y1, fp_sub(W1p, W2p));
  const fe X3pc = fp_sub(fp_sub(Dp, W1p), W2p);
  const fe e3 = fp_sub(X3pc, W1p);
  const fe C = SQR<CALLS>(e3);
  const fe A2 = fp_shl1(A1p);
  const fe Y3p = fp_sub(fp_sub(fp_sub(SQR<CALLS>(fp_add(dy, fp_sub(W1p, X3pc))), Dp), C), A2);
  const fe W1 = MUL<CALLS>(fp_shl<2>(X3pc), C);
  const fe W2 = MUL<CALLS>(fp_shl<2>(W1p), C);
  const fe ym = fp_sub(Y3p, A2);
  const fe yp = fp_add(Y3p, A2);
  const fe D = SQR<CALLS>(ym);
  const fe A1 = MUL<CALLS>(Y3p, fp_sub(W1, W2));
  const fe X3 = fp_sub(fp_sub(D, W1), W2);
  const fe Y3 = fp_sub(MUL<CALLS>(ym, fp_sub(W1, X3)), A1);
  const fe Z3 = MUL<CALLS>(Z, fp_sub(fp_sub(SQR<CALLS>(fp_sub(fp_add(dx, X3pc), W1p)), Cp), C));
  const fe Dc = SQR<CALLS>(yp);
  const fe X2n = fp_sub(fp_sub(Dc, W1), W2);
  const fe Y2n = fp_sub(MUL<CALLS>(yp, fp_sub(W1, X2n)), A1);
  X1 = X3; Y1 = Y3; X2 = X2n; Y2 = Y2n; Z = Z3;
}

// two phase groups per sub-partition, half a step apart (lockstep vs pipe mixing)
template <int THREADS>
__global__ void __launch_bounds__(THREADS, 1) k_zdau_skew2(uint32_t* io, int iters) {
  const size_t t = (size_t)threadIdx.x + (size_t)blockIdx.x * blockDim.x;
  const int grp = ((threadIdx.x >> 5) >> 2) & 1;
  fe v[5];
  Lazy md;
  for (int c = 0; c < 5; c++)
    for (int i = 0; i < 8; i++) v[c].v[i] = io[t * 40 + c * 8 + i];
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
    const uint32_t sw = (v[0].v[0] >> (it & 31)) & 1u;
    fe_cswap(sw, v[0], v[2]);
    fe_cswap(sw, v[1], v[3]);
    if (grp == 0) asm volatile("bar.sync 0;" ::: "memory");
    pt_zdau_xy<BENCH_QUIRK, Lazy, 1>(v[0], v[1], v[2], v[3], v[4], md, grp);
  }
  if (md.flagged()) v[0].v[0] ^= 1;
  for (int c = 0; c < 5; c++)
    for (int i = 0; i < 8; i++) io[t * 40 + c * 8 + i] = v[c].v[i];
}

#ifndef BENCH_UNROLL
#define BENCH_UNROLL 1
#endif
template <bool CALLS, bool SYNC, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k_zdau(uint32_t* io, int iters) {
  const size_t t = (size_t)threadIdx.x + (size_t)blockIdx.x * blockDim.x;
  fe v[5];
  Lazy md;
  for (int c = 0; c < 5; c++)
    for (int i = 0; i < 8; i++) v[c].v[i] = io[t * 40 + c * 8 + i];
  constexpr int kUnroll = BENCH_UNROLL;
#pragma unroll kUnroll
  for (int it = 0; it < iters; it++) {
    const uint32_t sw = (v[0].v[0] >> (it & 31)) & 1u;  // data-dependent swap like the ladder
    fe_cswap(sw, v[0], v[2]);
    fe_cswap(sw, v[1], v[3]);
    if (CALLS) zdau_v<CALLS>(v[0], v[1], v[2], v[3], v[4]);
    else pt_zdau_xy<BENCH_QUIRK>(v[0], v[1], v[2], v[3], v[4], md);
    if (SYNC) __syncthreads();
  }
  if (md.flagged()) v[0].v[0] ^= 1;  // keep the flags alive
  for (int c = 0; c < 5; c++)
    for (int i = 0; i < 8; i++) io[t * 40 + c * 8 + i] = v[c].v[i];
}

template <class K>
static void run(const char* name, K kern, int threads, int blocks_per_sm, uint32_t* d, int iters) {
  const int blocks = 148 * blocks_per_sm;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  kern<<<blocks, threads>>>(d, 4);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  kern<<<blocks, threads>>>(d, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaError_t err = cudaGetLastError();
  const double lanes = (double)blocks * threads;
  const double zdau_per_s = lanes * iters / (ms * 1e-3);
  const double clk_per_warp_iter = ms * 1e-3 * 1.965e9 * 148 * 4 / (lanes / 32 * iters);
  printf("{\"variant\": \"%s\", \"threads\": %d, \"blocks_per_sm\": %d, \"ms\": %.3f, \"zdau_per_s\": %.4g, \"scalar_mult_equiv_per_s\": %.4g, "
         "\"clk_per_warp_step_per_smsp\": %.0f, \"TMAC32_per_s\": %.3f, \"err\": \"%s\"}\n",
         name, threads, blocks_per_sm, ms, zdau_per_s, zdau_per_s / 255.5, clk_per_warp_iter, zdau_per_s * 828 / 1e12, cudaGetErrorString(err));
}

int main(int argc, char** argv) {
  const int iters = argc > 1 ? atoi(argv[1]) : 254;
  const size_t maxlanes = 148 * 2048;
  std::vector<uint32_t> h(maxlanes * 40);
  uint64_t s = 88172645463325252ull;
  for (auto& x : h) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; x = (uint32_t)s; }
  for (size_t i = 0; i < maxlanes * 5; i++) h[i * 8 + 7] &= 0x7fffffffu;  // canonical
  uint32_t* d;
  cudaMalloc(&d, h.size() * 4);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  run("inline_128x3", k_zdau<false, false, 128, 3>, 128, 3, d, iters);
  run("inline_128x4", k_zdau<false, false, 128, 4>, 128, 4, d, iters);
  run("inline_nosync_512x1", k_zdau<false, false, 512, 1>, 512, 1, d, iters);
  run("inline_sync_384x1", k_zdau<false, true, 384, 1>, 384, 1, d, iters);
  run("skew2_512x1", k_zdau_skew2<512>, 512, 1, d, iters);
  run("inline_sync_512x1", k_zdau<false, true, 512, 1>, 512, 1, d, iters);
  run("inline_sync_256x2", k_zdau<false, true, 256, 2>, 256, 2, d, iters);

#if VARIANT_CALLS
  run("calls_128x3", k_zdau<true, false, 128, 3>, 128, 3, d, iters);
  run("calls_sync_384x1", k_zdau<true, true, 384, 1>, 384, 1, d, iters);
  run("calls_128x4", k_zdau<true, false, 128, 4>, 128, 4, d, iters);
  run("calls_sync_512x1", k_zdau<true, true, 512, 1>, 512, 1, d, iters);
#endif
  return 0;
}
