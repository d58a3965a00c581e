// tools/pipe_probe.cu -- how do the FMA-heavy pipe (IMAD.WIDE) and the ALU pipe (SHF/LOP3/IADD3.X)
// share one SM sub-partition when the warps' instruction streams alternate between the two in
// bursts of B instructions?  (Development probe for DESIGN.md section 4.1; not part of the library.)
//   k_burst<B, KIND>: per trip 64 IMAD.WIDE + 128 ALU ops, as 64/B rounds of [B x IMAD.WIDE][2B x ALU];
//                     each round's ALU ops consume the round's products and feed the next round's
//                     multiplicands, so that ptxas cannot mix the two bursts.
//   k_indep<KIND>   : the same instruction counts as two independent streams (ptxas interleaves them).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/pipe_probe tools/pipe_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>

template <int KIND>
__device__ __forceinline__ void alu_op(uint32_t& x, uint32_t y) {
  if (KIND == 0) asm volatile("shf.l.wrap.b32 %0, %0, %1, 5;" : "+r"(x) : "r"(y));
  else if (KIND == 1) asm volatile("lop3.b32 %0, %0, %1, 0x5a5a5a5a, 0x96;" : "+r"(x) : "r"(y));
  else asm volatile("add.cc.u32 %0, %0, %1; addc.cc.u32 %0, %0, %1; addc.cc.u32 %0, %0, %1; addc.u32 %0, %0, %1;" : "+r"(x) : "r"(y));
}
__host__ __device__ constexpr int alu_per_call(int KIND) { return KIND == 2 ? 4 : 1; }

__device__ __forceinline__ void wide(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
  asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
}

template <int B, int KIND>
__global__ void __launch_bounds__(512, 1) k_burst(uint32_t* out, const uint32_t* in, int iters) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t lo[8], hi[8], x[8];
#pragma unroll
  for (int j = 0; j < 8; j++) { lo[j] = in[j] + tid; hi[j] = ~lo[j]; x[j] = lo[j] * 3u + 1u; }
  const uint32_t y = in[8];
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 64 / B; r++) {
#pragma unroll
      for (int j = 0; j < B; j++) wide(lo[j & 7], hi[j & 7], x[(j + r) & 7], y);
#pragma unroll
      for (int j = 0; j < 2 * B / alu_per_call(KIND); j++) alu_op<KIND>(x[j & 7], hi[(j + 3) & 7]);
    }
  }
  uint32_t t = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) t += lo[j] ^ hi[j] ^ x[j];
  out[tid] = t;
}

template <int KIND, int NW, int NA>
__global__ void __launch_bounds__(512, 1) k_indep(uint32_t* out, const uint32_t* in, int iters) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t lo[8], hi[8], x[8];
#pragma unroll
  for (int j = 0; j < 8; j++) { lo[j] = in[j] + tid; hi[j] = ~lo[j]; x[j] = lo[j] * 3u + 1u; }
  const uint32_t y = in[8];
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int j = 0; j < 64; j++) {
      if (j < NW) wide(lo[j & 7], hi[j & 7], hi[(j + 3) & 7], y);
      if (2 * j < NA) { alu_op<KIND>(x[(2 * j) & 7], x[(2 * j + 3) & 7]); if (KIND != 2) alu_op<KIND>(x[(2 * j + 1) & 7], x[(2 * j + 4) & 7]); }
    }
    if (KIND == 2) { /* 4 adds per call: 64 calls made above give 256; trim is not needed for the comparison */ }
  }
  uint32_t t = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) t += lo[j] ^ hi[j] ^ x[j];
  out[tid] = t;
}

// warp-specialised: warps 0..(n/2-1) of every sub-partition run only IMAD.WIDE, the others only ALU ops
template <int KIND>
__global__ void __launch_bounds__(512, 1) k_spec(uint32_t* out, const uint32_t* in, int iters) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  const bool wwarp = ((threadIdx.x >> 5) >> 2) & 1;  // warps 4-7, 12-15 multiply: two of each kind per sub-partition
  uint32_t lo[8], hi[8], x[8];
#pragma unroll
  for (int j = 0; j < 8; j++) { lo[j] = in[j] + tid; hi[j] = ~lo[j]; x[j] = lo[j] * 3u + 1u; }
  const uint32_t y = in[8];
  if (wwarp) {
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int j = 0; j < 128; j++) wide(lo[j & 7], hi[j & 7], hi[(j + 3) & 7], y);
    }
  } else {
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int j = 0; j < 256 / alu_per_call(KIND); j++) alu_op<KIND>(x[j & 7], x[(j + 3) & 7]);
    }
  }
  uint32_t t = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) t += lo[j] ^ hi[j] ^ x[j];
  out[tid] = t;
}

template <class K>
static void run(const char* name, K kern, int threads, uint32_t* dout, uint32_t* din, int iters) {
  const int blocks = 148;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  kern<<<blocks, threads>>>(dout, din, 8);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int rep = 0; rep < 3; rep++) {
    cudaEventRecord(e0);
    kern<<<blocks, threads>>>(dout, din, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  const double clk = best * 1e-3 * 1.965e9 / iters;  // per trip, all warps of a sub-partition together
  const int wps = threads / 128;
  printf("{\"probe\": \"%s\", \"warps_per_smsp\": %d, \"clk_per_trip_smsp\": %.1f, \"clk_per_warp_trip\": %.1f, \"err\": \"%s\"}\n", name, wps, clk,
         clk / wps, cudaGetErrorString(cudaGetLastError()));
}

int main(int argc, char** argv) {
  const int iters = argc > 1 ? atoi(argv[1]) : 2000;
  uint32_t *dout, *din;
  cudaMalloc(&dout, 148 * 1024 * 4);
  cudaMalloc(&din, 64);
  uint32_t h[16] = {1, 2, 3, 4, 5, 6, 7, 8, 0x9e3779b9u};
  cudaMemcpy(din, h, 64, cudaMemcpyHostToDevice);
  for (int threads : {128, 256, 512, 1024}) {
#define RUNB(B, K) run("burst" #B "_kind" #K, k_burst<B, K>, threads, dout, din, iters)
    RUNB(64, 0); RUNB(16, 0); RUNB(4, 0); RUNB(1, 0);
    RUNB(64, 1); RUNB(4, 1);
    RUNB(64, 2); RUNB(16, 2); RUNB(4, 2);
    run("indep_kind0", k_indep<0, 64, 128>, threads, dout, din, iters);
    run("indep_kind1", k_indep<1, 64, 128>, threads, dout, din, iters);
    run("indep_kind2_256adds", k_indep<2, 64, 128>, threads, dout, din, iters);
    run("only_wide64", k_indep<0, 64, 0>, threads, dout, din, iters);
    run("only_alu128_kind0", k_indep<0, 0, 128>, threads, dout, din, iters);
    run("only_alu128_kind1", k_indep<1, 0, 128>, threads, dout, din, iters);
    run("only_alu256_kind2", k_indep<2, 0, 128>, threads, dout, din, iters);
    if (threads >= 256) {
      run("spec_kind0_(128W|256A per pair)", k_spec<0>, threads, dout, din, iters);
      run("spec_kind1", k_spec<1>, threads, dout, din, iters);
      run("spec_kind2", k_spec<2>, threads, dout, din, iters);
    }
  }
  return 0;
}
