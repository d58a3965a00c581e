// tools/pipe_probe4.cu -- what does an ALU instruction cost next to IMAD.WIDE as a function of its REGISTER READS?
// 64 conflict-free IMAD.WIDE (even * odd multiplicands, see pipe_probe3.cu) per trip on 8 accumulator pairs, plus
// NA ALU instructions that work on halves of the same pairs, so their operand banks are known (low word = even
// register, high word = odd register; checked in the SASS):
//   K0  lop3  lo ^= hi'          2 register reads, one per bank
//   K1  lop3  lo ^= lo'          2 reads, same bank (even)
//   K2  lop3  lo ^= imm          1 read
//   K3  lop3  lo = f(lo, hi', lo'')   3 reads
//   K4  lop3  hi ^= hi'          2 reads, same bank (odd)
//   K5  mov   lo = imm           0 reads          (kept alive by the next multiply)
//   K6  ffma2 pair = pair * scalar(broadcast) + imm   3 reads for two FMAs   (on a separate fp32 pair set)
//   K7  ffma  x 2 (the two FMAs K6 replaces)          2 reads each
#include <cstdio>
#include <cstdlib>
#include <cstdint>

template <int K, int NW, int NA>
__global__ void __launch_bounds__(512, 1) k_mix4(uint32_t* out, const uint2* in, int iters) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t lo[8], hi[8], x[8];
  unsigned long long f2[8];
  float f[8], g[8];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const uint2 X = in[(tid & 31) + 32 * j];
    lo[j] = X.x * 3u + tid; hi[j] = X.y ^ tid; x[j] = X.x ^ (X.y >> 3);
    f[j] = 1.0f + (float)j * 0.01f + (float)(tid & 7) * 0.001f; g[j] = 1.5f - (float)j * 0.01f;
    f2[j] = ((unsigned long long)__float_as_uint(g[j]) << 32) | __float_as_uint(f[j]);
  }
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int gq = 0; gq < 64; gq++) {
#pragma unroll
      for (int k = (gq * NW) / 64; k < ((gq + 1) * NW) / 64; k++) {
        const int i = (k + 3) & 7, j0 = (k * 5 + 1) & 7, j = (j0 == (k & 7)) ? ((j0 + 1) & 7) : j0;
        asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(lo[k & 7]), "+r"(hi[k & 7]) : "r"(lo[i]), "r"(hi[j]));
      }
#pragma unroll
      for (int k = (gq * NA) / 64; k < ((gq + 1) * NA) / 64; k++) {
        const int d = k & 7, s = (k + 3) & 7, s2 = (k + 5) & 7;
        if (K == 0) asm volatile("lop3.b32 %0, %0, %1, 0x5a5a5a5a, 0x96;" : "+r"(x[d]) : "r"(x[s]));
        if (K == 2) asm volatile("lop3.b32 %0, %0, 0x12345677, 0x5a5a5a5a, 0x96;" : "+r"(x[d]));
        if (K == 3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[d]) : "r"(x[s]), "r"(x[s2]));
        if (K == 6) asm volatile("{ .reg .b64 bb; mov.b64 bb, {%1, %1}; fma.rn.f32x2 %0, %0, bb, %2; }" : "+l"(f2[d]) : "f"(f[s]), "l"(0x3f8000003f800000ull));
        if (K == 7) { asm volatile("fma.rn.f32 %0, %0, %1, 1.0;" : "+f"(g[d]) : "f"(f[s])); asm volatile("fma.rn.f32 %0, %0, %1, 1.0;" : "+f"(f[d]) : "f"(g[s])); }
        if (K == 8) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(f2[d]) : "l"(f2[s]), "l"(0x3f8000003f800000ull));
      }
    }
  }
  unsigned long long t = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) t += lo[j] ^ hi[j] ^ x[j] ^ f2[j] ^ __float_as_uint(f[j]) ^ __float_as_uint(g[j]);
  out[tid] = (uint32_t)t ^ (uint32_t)(t >> 32);
}

template <class Kn>
static void run(const char* name, Kn kern, int threads, uint32_t* dout, uint2* din, int iters) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  kern<<<148, threads>>>(dout, din, 8);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int rep = 0; rep < 3; rep++) {
    cudaEventRecord(e0);
    kern<<<148, threads>>>(dout, din, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  const double clk = best * 1e-3 * 1.965e9 / iters;
  const int wps = threads / 128;
  printf("{\"probe\": \"%s\", \"warps_per_smsp\": %d, \"clk_per_warp_trip\": %.1f, \"err\": \"%s\"}\n", name, wps, clk / wps, cudaGetErrorString(cudaGetLastError()));
}

int main(int argc, char** argv) {
  const int iters = argc > 1 ? atoi(argv[1]) : 2000;
  uint32_t* dout;
  uint2* din;
  cudaMalloc(&dout, 148 * 1024 * 4);
  cudaMalloc(&din, 32 * 16 * 8);
  uint2 h[32 * 16];
  for (int i = 0; i < 32 * 16; i++) h[i] = make_uint2(0x9e3779b9u * (i + 1), 0x85ebca6bu * (i + 3));
  cudaMemcpy(din, h, sizeof(h), cudaMemcpyHostToDevice);
  const int threads = 512;
#define RUN(K, NW, NA) run("k" #K "_W" #NW "_A" #NA, k_mix4<K, NW, NA>, threads, dout, din, iters)
  RUN(0, 64, 0);
  RUN(0, 64, 128); RUN(2, 64, 128); RUN(3, 64, 128); RUN(6, 64, 64); RUN(7, 64, 64); RUN(8, 64, 64);
  RUN(0, 0, 128); RUN(2, 0, 128); RUN(3, 0, 128); RUN(6, 0, 64); RUN(7, 0, 64); RUN(8, 0, 64);
  RUN(0, 64, 256); RUN(2, 64, 256); RUN(3, 64, 256);
  return 0;
}
