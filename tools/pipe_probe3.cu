// tools/pipe_probe3.cu -- round-2 probes of the sub-partition issue port (see pipe_probe2.cu):
//   (1) does the register-bank placement of the IMAD.WIDE operands matter?  The multiplicands come out of
//       64-bit loads (aligned pairs: .x in an even register, .y in an odd one), the addend is an aligned
//       pair (one word in each bank), so  even*odd  reads two registers per bank and  even*even / odd*odd
//       reads three from one bank.  SASS checked with cuobjdump (tools/README.md).
//   (2) what does FFMA2 (fma.rn.f32x2, two fp32 FMAs per instruction) cost next to IMAD.WIDE, compared
//       with the two FFMAs it replaces?
// per trip: NW wide multiplies (bank mode BK), NF FFMAs, NF2 FFMA2s, NA LOP3s spread evenly over 64 groups.
#include <cstdio>
#include <cstdlib>
#include <cstdint>

// BK: 0 = even*odd, 1 = even*even, 2 = odd*odd, 3 = even*odd with the same multiplicand for 8 in a row,
//     4 = the same register as both multiplicands; 3/5/6 take one multiplicand from a register that never changes (X[i].x, even):
//     3 = even*odd, the same X for 8 in a row; 5 = even*even, the same X for 8 in a row (reuse cache?); 6 = even*even, another X each time
template <int BK, int NW, int NF, int NF2, int NA>
__global__ void __launch_bounds__(512, 1) k_mix3(uint32_t* out, const uint2* in, int iters) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  uint2 X[8], Y[8];
  uint32_t lo[8], hi[8];
  float f[8];
  unsigned long long f2[8];
  uint32_t x[8];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    X[j] = in[(tid & 31) + 32 * j];
    Y[j] = in[(tid & 31) + 32 * (j + 8)];
    lo[j] = X[j].x * Y[j].y + tid; hi[j] = X[j].y ^ Y[j].x;
    f[j] = (float)j + tid;
    f2[j] = ((unsigned long long)__float_as_uint(1.0f + j) << 32) | __float_as_uint(2.0f + tid);
    x[j] = X[j].x ^ Y[j].y;
  }
  const float fy = 1.0001f;
  const unsigned long long fy2 = ((unsigned long long)__float_as_uint(1.0001f) << 32) | __float_as_uint(0.9999f);
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int g = 0; g < 64; g++) {
#pragma unroll
      for (int k = (g * NW) / 64; k < ((g + 1) * NW) / 64; k++) {
        // multiplicands are words of OTHER accumulators (low word = even register, high word = odd register of an
        // aligned pair), so every product differs and ptxas can neither hoist nor merge them
        const int i = (BK == 3 || BK == 5) ? ((k >> 3) & 7) : ((k + 3) & 7), j = (k * 5 + 1) & 7;
        const int jj = (j == (k & 7)) ? ((j + 1) & 7) : j;
        const uint32_t a = (BK == 3 || BK == 5 || BK == 6) ? X[i].x : (BK == 2) ? hi[i] : lo[i];
        const uint32_t b = (BK == 4) ? a : (BK == 1 || BK == 5 || BK == 6) ? lo[jj] : hi[jj];
        asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(lo[k & 7]), "+r"(hi[k & 7]) : "r"(a), "r"(b));
      }
#pragma unroll
      for (int k = (g * NF) / 64; k < ((g + 1) * NF) / 64; k++)
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[k & 7]) : "f"(fy), "f"(f[(k + 3) & 7]));
#pragma unroll
      for (int k = (g * NF2) / 64; k < ((g + 1) * NF2) / 64; k++)
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(f2[k & 7]) : "l"(fy2), "l"(f2[(k + 3) & 7]));
#pragma unroll
      for (int k = (g * NA) / 64; k < ((g + 1) * NA) / 64; k++)
        asm volatile("lop3.b32 %0, %0, %1, 0x5a5a5a5a, 0x96;" : "+r"(x[k & 7]) : "r"(x[(k + 3) & 7]));
    }
  }
  unsigned long long t = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) t += lo[j] ^ hi[j] ^ f2[j] ^ __float_as_uint(f[j]) ^ x[j] ^ X[j].y ^ Y[j].x;
  out[tid] = (uint32_t)t ^ (uint32_t)(t >> 32);
}

template <class K>
static void run(const char* name, K kern, int threads, uint32_t* dout, uint2* din, int iters) {
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
  for (int threads : {256, 512}) {
#define RUN(BK, NW, NF, NF2, NA) run("bk" #BK "_W" #NW "_F" #NF "_FF" #NF2 "_A" #NA, k_mix3<BK, NW, NF, NF2, NA>, threads, dout, din, iters)
    RUN(0, 64, 0, 0, 0); RUN(1, 64, 0, 0, 0); RUN(2, 64, 0, 0, 0); RUN(3, 64, 0, 0, 0); RUN(4, 64, 0, 0, 0); RUN(5, 64, 0, 0, 0); RUN(6, 64, 0, 0, 0);
    RUN(0, 64, 0, 0, 128); RUN(1, 64, 0, 0, 128); RUN(2, 64, 0, 0, 128); RUN(3, 64, 0, 0, 128);
    RUN(0, 0, 128, 0, 0); RUN(0, 0, 0, 64, 0); RUN(0, 0, 0, 128, 0);
    RUN(0, 64, 128, 0, 0); RUN(0, 64, 0, 64, 0);
    RUN(0, 64, 64, 0, 128); RUN(0, 64, 0, 32, 128);
    RUN(0, 0, 0, 64, 128); RUN(0, 0, 128, 0, 128);
  }
  return 0;
}
