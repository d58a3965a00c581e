// host_common.cuh -- host-side plumbing shared by the C-ABI translation units:
// error reporting, stream-ordered temporaries, host<->device staging.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <vector>

#include "../../include/ecb200.h"

namespace ecb200 {

void set_error(const char* fmt, ...);
const char* last_error();
extern std::atomic<uint64_t> g_launches;

// Per-device state owned by the library (kernels_field.cu): its own stream-ordered memory pool (the
// process-wide default pool is left alone), the fixed-base tables, the three pipeline streams of the
// host-memory batch path with their pinned bounce buffers.  Created by ecb200_init / ecb200_init_devices
// (or lazily by the first call on a device), released by ecb200_shutdown.
constexpr int kMaxDevices = 64;
struct DeviceCtx {
  std::mutex mu;                        // guards the lazily created members below
  bool ready = false;
  int device = -1;
  cudaMemPool_t pool = nullptr;
  uint4* base_tab[2] = {nullptr, nullptr};   // [quirk] 2^16 ladder states of G
  std::mutex pipe_mu;                   // one host-memory batch at a time per device
  cudaStream_t pipe[3] = {nullptr, nullptr, nullptr};
  void* bounce_in[3] = {nullptr, nullptr, nullptr};   // pinned staging for pageable caller buffers
  void* bounce_out[3] = {nullptr, nullptr, nullptr};
  size_t bounce_in_bytes = 0, bounce_out_bytes = 0;
  void* slot_dev[3] = {nullptr, nullptr, nullptr};    // device staging of one chunk per pipeline stream (k, P, result, affine result)
};
// context of the calling thread's current device, created on first use; nullptr + error set on failure
DeviceCtx* current_ctx();

#define ECB_CUDA(expr)                                                                   \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      ::ecb200::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return ECB200_ERR_CUDA;                                                            \
    }                                                                                    \
  } while (0)

#define ECB_LAUNCH_CHECK()                                                               \
  do {                                                                                   \
    ::ecb200::g_launches.fetch_add(1, std::memory_order_relaxed);                        \
    cudaError_t _e = cudaGetLastError();                                                 \
    if (_e != cudaSuccess) {                                                             \
      ::ecb200::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return ECB200_ERR_CUDA;                                                            \
    }                                                                                    \
  } while (0)

inline int layout_of(uint32_t flags) { return (int)(flags & ECB200_LAYOUT_MASK); }
inline bool on_device(uint32_t flags) { return (flags & ECB200_MEM_MASK) == ECB200_MEM_DEVICE; }
inline bool quirk_on(uint32_t flags) { return (flags & ECB200_NO_QUIRK) == 0; }

inline int check_common(size_t n, uint32_t flags) {
  const int L = layout_of(flags);
  if (L != (int)ECB200_LAYOUT_LANE && L != (int)ECB200_LAYOUT_PACK4 && L != (int)ECB200_LAYOUT_SOA) {
    set_error("unknown layout %d", L);
    return ECB200_ERR_ARG;
  }
  if (L == (int)ECB200_LAYOUT_PACK4 && (n & 3)) {
    set_error("PACK4 layout needs n %% 4 == 0 (n = %zu)", n);
    return ECB200_ERR_ARG;
  }
  return ECB200_OK;
}

// One call's worth of device temporaries (stream-ordered; freed on the same stream).
struct Scratch {
  cudaStream_t stream;
  DeviceCtx* ctx = nullptr;
  std::vector<void*> ptrs;
  explicit Scratch(cudaStream_t s) : stream(s) {}
  ~Scratch() {
    for (void* p : ptrs) cudaFreeAsync(p, stream);
  }
  int alloc(void** out, size_t bytes) {
    if (bytes == 0) bytes = 16;
    if (!ctx && !(ctx = current_ctx())) return ECB200_ERR_CUDA;
    cudaError_t e = cudaMallocFromPoolAsync(out, bytes, ctx->pool, stream);
    if (e != cudaSuccess) {
      set_error("cudaMallocFromPoolAsync(%zu) failed: %s", bytes, cudaGetErrorString(e));
      return e == cudaErrorMemoryAllocation ? ECB200_ERR_NOMEM : ECB200_ERR_CUDA;
    }
    ptrs.push_back(*out);
    return ECB200_OK;
  }
};

// An operand of a batched call: pointer + number of 256-bit coordinates per lane.
struct Operand {
  const void* in;   // input pointer (caller space) or nullptr
  void* out;        // output pointer (caller space) or nullptr
  int nc;
  void* dev;        // device pointer the kernel uses (filled by stage_*)
};

inline size_t operand_bytes(size_t n, int nc) { return n * (size_t)nc * 32; }

}  // namespace ecb200
