// tools/variant_bench.cu -- times the stand-alone ladder kernel (the TU of tools/order_search.py / tools/ladder_census.sh,
// entry point k_ladder, SOA layout) of every .cubin given on the command line, so that many compile variants can be
// measured in one GPU call.  Inputs are pseudo-random words (the ladder's time does not depend on the data; parity is
// the test suite's job, not this tool's); every variant also prints a checksum of its output, which must agree
// between variants that compute the same thing.
// Build: nvcc -O2 -std=c++17 -o build/variant_bench tools/variant_bench.cu -lcuda
// Run:   build/variant_bench <log2 lanes> a.cubin b.cubin ...
// With KERNEL=<mangled name> in the environment the cubins are variants of the library's own translation unit
// (kernels_point.cubin) and that instance of k_scalar_mult_sync is launched (out, k, P, n, k_bcast = 0, tab = NULL):
// tools/recolor_autotune.py times several re-colourings of a shipped kernel this way.
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

#define CK(x) do { CUresult r_ = (x); if (r_ != CUDA_SUCCESS) { const char* s_; cuGetErrorString(r_, &s_); printf("{\"error\": \"%s at %s:%d\"}\n", s_, __FILE__, __LINE__); exit(1); } } while (0)

static uint64_t splitmix(uint64_t& s) { uint64_t z = (s += 0x9e3779b97f4a7c15ull); z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull; z = (z ^ (z >> 27)) * 0x94d049bb133111ebull; return z ^ (z >> 31); }

int main(int argc, char** argv) {
  if (argc < 3) { printf("usage: %s log2n cubin...\n", argv[0]); return 2; }
  const size_t n = (size_t)1 << atoi(argv[1]);
  CK(cuInit(0));
  CUdevice dev; CK(cuDeviceGet(&dev, 0));
  CUcontext ctx; CK(cuCtxCreate(&ctx, 0, dev));
  CUdeviceptr dk, dP, dout;
  CK(cuMemAlloc(&dk, n * 32)); CK(cuMemAlloc(&dP, n * 96)); CK(cuMemAlloc(&dout, n * 96));
  std::vector<uint32_t> h(n * 24);
  uint64_t s = 0xEC51D003;
  for (auto& w : h) w = (uint32_t)splitmix(s) & 0x7fffffffu;   // top limb below p: canonical-looking operands
  CK(cuMemcpyHtoD(dP, h.data(), n * 96));
  for (size_t i = 0; i < n * 8; i++) h[i] = (uint32_t)splitmix(s);
  CK(cuMemcpyHtoD(dk, h.data(), n * 32));
  CUevent e0, e1; CK(cuEventCreate(&e0, 0)); CK(cuEventCreate(&e1, 0));
  for (int a = 2; a < argc; a++) {
    CUmodule mod;
    if (cuModuleLoad(&mod, argv[a]) != CUDA_SUCCESS) { printf("{\"variant\": \"%s\", \"error\": \"load\"}\n", argv[a]); continue; }
    const char* kname = getenv("KERNEL");
    CUfunction fn; CK(cuModuleGetFunction(&fn, mod, kname ? kname : "k_ladder"));
    int threads = 512;
    CK(cuFuncGetAttribute(&threads, CU_FUNC_ATTRIBUTE_MAX_THREADS_PER_BLOCK, fn));
    if (threads > 512) threads = 512;
    int regs = 0; CK(cuFuncGetAttribute(&regs, CU_FUNC_ATTRIBUTE_NUM_REGS, fn));
    const unsigned smem = 5 * 2 * 512 * 16;   // LadderSmem<512>::kBytes (unused by the register-state variants)
    CK(cuFuncSetAttribute(fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, smem));
    size_t nn = n;
    int k_bcast = 0;
    CUdeviceptr tab = 0;
    if (const char* tk = getenv("TABLE_KERNEL")) {   // fixed-base instance: build the 2^16-entry table with the module's own kernel
      CUfunction tf; CK(cuModuleGetFunction(&tf, mod, tk));
      CK(cuMemAlloc(&tab, (size_t)160 << 16));
      int tabw = 16;
      void* targs[] = {&tab, &tabw};
      CK(cuLaunchKernel(tf, (1u << 16) / 128, 1, 1, 128, 1, 1, 0, 0, targs, 0));
      CK(cuCtxSynchronize());
    }
    void* args[] = {&dout, &dk, &dP, &nn, &k_bcast, &tab};   // the stand-alone k_ladder takes the first four
    const unsigned blocks = (unsigned)((n + threads - 1) / threads);
    CK(cuMemsetD8(dout, 0, n * 96));
    CK(cuLaunchKernel(fn, blocks, 1, 1, threads, 1, 1, smem, 0, args, 0));
    CK(cuCtxSynchronize());
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
      CK(cuEventRecord(e0, 0));
      CK(cuLaunchKernel(fn, blocks, 1, 1, threads, 1, 1, smem, 0, args, 0));
      CK(cuEventRecord(e1, 0));
      CK(cuEventSynchronize(e1));
      float ms; CK(cuEventElapsedTime(&ms, e0, e1));
      if (ms < best) best = ms;
    }
    std::vector<uint32_t> o(n * 24);
    uint64_t sum = 0;
    int unstable = 0;   // REPEAT=<r> in the environment: r more runs whose outputs must all have the same checksum
    const int repeat = getenv("REPEAT") ? atoi(getenv("REPEAT")) : 0;
    for (int rep = 0; rep <= repeat; rep++) {
      if (rep) { CK(cuLaunchKernel(fn, blocks, 1, 1, threads, 1, 1, smem, 0, args, 0)); }
      CK(cuMemcpyDtoH(o.data(), dout, n * 96));
      uint64_t s2 = 0;
      for (size_t i = 0; i < o.size(); i++) s2 = s2 * 1099511628211ull + o[i];
      if (rep && s2 != sum) unstable++;
      if (!rep) sum = s2;
    }
    if (tab) CK(cuMemFree(tab));
    printf("{\"variant\": \"%s\", \"lanes\": %zu, \"threads\": %d, \"regs\": %d, \"ms\": %.4f, \"Mps\": %.3f, \"checksum\": \"%016llx\", \"unstable_runs\": %d}\n", argv[a], n, threads, regs, best,
           n / best * 1e-3, (unsigned long long)sum, unstable);
    fflush(stdout);
    CK(cuModuleUnload(mod));
  }
  return 0;
}
