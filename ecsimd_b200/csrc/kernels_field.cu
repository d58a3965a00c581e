// kernels_field.cu -- streaming Montgomery field kernels (one lane per thread), the
// layout conversion kernels, synthetic input generation, and the library/runtime part
// of the C ABI declared in include/ecb200.h.
#include <mutex>

#include "host_common.cuh"
#include "layout.cuh"

namespace ecb200 {

static thread_local char t_err[512] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_err, sizeof t_err, fmt, ap);
  va_end(ap);
}
const char* last_error() { return t_err; }

// ---- per-device contexts ---------------------------------------------------------------------------
static DeviceCtx g_ctx[kMaxDevices];

static int ensure_ctx(int device, DeviceCtx** out) {
  if (device < 0 || device >= kMaxDevices) {
    set_error("device %d is outside the %d contexts this library keeps", device, kMaxDevices);
    return ECB200_ERR_ARG;
  }
  DeviceCtx& c = g_ctx[device];
  std::lock_guard<std::mutex> lock(c.mu);
  if (!c.ready) {
    int major = 0, minor = 0;
    ECB_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
    ECB_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device));
    if (major != 10) {
      set_error("device %d is sm_%d%d; this library carries sm_100a code only", device, major, minor);
      return ECB200_ERR_CUDA;
    }
    // the library's own pool for its stream-ordered temporaries, kept warm between calls; the process-wide
    // default pool of the device is not touched
    cudaMemPoolProps props;
    memset(&props, 0, sizeof props);
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = device;
    ECB_CUDA(cudaMemPoolCreate(&c.pool, &props));
    unsigned long long thr = ~0ull;
    ECB_CUDA(cudaMemPoolSetAttribute(c.pool, cudaMemPoolAttrReleaseThreshold, &thr));
    c.device = device;
    c.ready = true;
  }
  *out = &c;
  return ECB200_OK;
}

DeviceCtx* current_ctx() {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    set_error("cudaGetDevice failed: %s", cudaGetErrorString(e));
    return nullptr;
  }
  DeviceCtx* c = nullptr;
  return ensure_ctx(dev, &c) == ECB200_OK ? c : nullptr;
}

// devices a host-memory batch is cut over (ecb200_init_devices); empty = the calling thread's device only
static std::mutex g_multi_mu;
static std::vector<int> g_multi;
std::vector<int> multi_devices() {
  std::lock_guard<std::mutex> lock(g_multi_mu);
  return g_multi;
}

enum FieldOp : int { OP_ADD, OP_SUB, OP_MUL, OP_SQR, OP_SHL, OP_NEG, OP_FROMC, OP_TOC, OP_INV, OP_MULCHAIN };

// The inversion runs in Lazy mode (fp_pow_lsb, fp256.cuh); a lane that met one of the 2^-32 corner cases is
// recomputed from its input in Exact mode, out of line.
template <bool QUIRK>
__device__ __noinline__ void fp_inv_exact(uint32_t* r8, const uint32_t* a8) {
  Exact md;
  fe a;
  for (int i = 0; i < 8; i++) a.v[i] = a8[i];
  const fe r = fp_inv<QUIRK>(a, md);
  for (int i = 0; i < 8; i++) r8[i] = r.v[i];
}
template <bool QUIRK>
__device__ __forceinline__ fe fp_inv(const fe& a) {
  Lazy md;
  fe r = fp_inv<QUIRK>(a, md);
  if (__builtin_expect(md.flagged(), 0)) {
    uint32_t in[8], out[8];
#pragma unroll
    for (int i = 0; i < 8; i++) in[i] = a.v[i];
    fp_inv_exact<QUIRK>(out, in);
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = out[i];
  }
  return r;
}

template <int L, int OP, bool QUIRK>
__global__ void __launch_bounds__(256) k_field(void* __restrict__ out, const void* __restrict__ a,
                                               const void* __restrict__ b, size_t n, int param) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  fe x = Layout<L>::load(a, n, i, 1, 0);
  fe r;
  if (OP == OP_ADD) r = fp_add(x, Layout<L>::load(b, n, i, 1, 0));
  else if (OP == OP_SUB) r = fp_sub(x, Layout<L>::load(b, n, i, 1, 0));
  else if (OP == OP_MUL) r = fp_mul(x, Layout<L>::load(b, n, i, 1, 0));
  else if (OP == OP_SQR) r = fp_sqr<QUIRK>(x);
  else if (OP == OP_SHL) { r = x; for (int c = 0; c < param; c++) r = fp_shl1(r); }
  else if (OP == OP_NEG) r = fp_neg(x);
  else if (OP == OP_FROMC) r = fp_from_classical(x);
  else if (OP == OP_TOC) r = fp_to_classical(x);
  else if (OP == OP_INV) r = fp_inv<QUIRK>(x);
  else if (OP == OP_MULCHAIN) {
    const fe y = Layout<L>::load(b, n, i, 1, 0);
    r = x;
#pragma unroll 4
    for (int c = 0; c < param; c++) r = fp_mul(r, y);
  }
  Layout<L>::store(out, n, i, 1, 0, r);
}

template <int L>
__global__ void __launch_bounds__(256) k_to_soa(void* __restrict__ dst, const void* __restrict__ src, size_t n, int nc) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int c = 0; c < nc; c++) Layout<L_SOA>::store(dst, n, i, nc, c, Layout<L>::load(src, n, i, nc, c));
}
template <int L>
__global__ void __launch_bounds__(256) k_from_soa(void* __restrict__ dst, const void* __restrict__ src, size_t n, int nc) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int c = 0; c < nc; c++) Layout<L>::store(dst, n, i, nc, c, Layout<L_SOA>::load(src, n, i, nc, c));
}

template <int LS, int LD>
__global__ void __launch_bounds__(256) k_convert(void* __restrict__ dst, const void* __restrict__ src, size_t n, int nc) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int c = 0; c < nc; c++) Layout<LD>::store(dst, n, i, nc, c, Layout<LS>::load(src, n, i, nc, c));
}

// serialization.h:12-48 of the reference: a value <-> its 32-byte big-endian string.  `bytes` holds
// n*nc strings back to back (lane-major, coordinate-minor: SEC1-style x|y for nc = 2).
template <int L, bool TO_BYTES>
__global__ void __launch_bounds__(256) k_bytes_be(void* __restrict__ vals, uint8_t* __restrict__ bytes, size_t n, int nc) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int c = 0; c < nc; c++) {
    uint4* b = reinterpret_cast<uint4*>(bytes) + (i * nc + c) * 2;
    if (TO_BYTES) {
      const fe v = Layout<L>::load(vals, n, i, nc, c);
      b[0] = make_uint4(__byte_perm(v.v[7], 0, 0x0123), __byte_perm(v.v[6], 0, 0x0123), __byte_perm(v.v[5], 0, 0x0123), __byte_perm(v.v[4], 0, 0x0123));
      b[1] = make_uint4(__byte_perm(v.v[3], 0, 0x0123), __byte_perm(v.v[2], 0, 0x0123), __byte_perm(v.v[1], 0, 0x0123), __byte_perm(v.v[0], 0, 0x0123));
    } else {
      const uint4 hi = b[0], lo = b[1];
      fe v;
      v.v[7] = __byte_perm(hi.x, 0, 0x0123); v.v[6] = __byte_perm(hi.y, 0, 0x0123); v.v[5] = __byte_perm(hi.z, 0, 0x0123); v.v[4] = __byte_perm(hi.w, 0, 0x0123);
      v.v[3] = __byte_perm(lo.x, 0, 0x0123); v.v[2] = __byte_perm(lo.y, 0, 0x0123); v.v[1] = __byte_perm(lo.z, 0, 0x0123); v.v[0] = __byte_perm(lo.w, 0, 0x0123);
      Layout<L>::store(vals, n, i, nc, c, v);
    }
  }
}

__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
  unsigned long long z = x + 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

template <int L>
__global__ void __launch_bounds__(256) k_synth(void* __restrict__ out, unsigned long long seed, unsigned long long start,
                                               int kind, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  fe r;
#pragma unroll
  for (int l = 0; l < 4; l++) {
    const unsigned long long w = splitmix64(seed * 0x100000001B3ull + 4ull * (start + i) + (unsigned long long)l);
    r.v[2 * l] = (uint32_t)w;
    r.v[2 * l + 1] = (uint32_t)(w >> 32);
  }
  if (kind == 1) {
    // canonical field element: subtract p once if >= p  (the carry-free form of fp_reduce_once)
    r = fp_reduce_once(r, 0);
  }
  Layout<L>::store(out, n, i, 1, 0, r);
}

__global__ void __launch_bounds__(256) k_checksum(uint32_t* __restrict__ acc8, const uint32_t* __restrict__ buf, size_t nwords) {
  // order-independent fold: word j contributes to slot (j & 7) by XOR and to slot 8+(j&7) by +
  uint32_t x = 0;
  const size_t stride = (size_t)gridDim.x * blockDim.x * 8;
  size_t j = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  uint32_t lane[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (; j + 8 <= nwords; j += stride) {
    const uint4 p = *reinterpret_cast<const uint4*>(buf + j);
    const uint4 q = *reinterpret_cast<const uint4*>(buf + j + 4);
    lane[0] ^= p.x; lane[1] ^= p.y; lane[2] ^= p.z; lane[3] ^= p.w;
    lane[4] ^= q.x; lane[5] ^= q.y; lane[6] ^= q.z; lane[7] ^= q.w;
  }
  (void)x;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    uint32_t v = lane[k];
    for (int o = 16; o > 0; o >>= 1) v ^= __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicXor(acc8 + k, v);
  }
}

// ---- integer-pipe micro-benchmarks --------------------------------------------------------
template <int WHICH>
__global__ void __launch_bounds__(256) k_microbench(uint32_t* __restrict__ out, const uint32_t* __restrict__ in, int iters) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t x[8], y = in[8];
#pragma unroll
  for (int j = 0; j < 8; j++) x[j] = in[j] + tid;
  // asm volatile keeps nvcc/ptxas from strength-reducing the loops
#define ECB_ADD8(S)                                                                                   \
  asm volatile("add.cc.u32 %0, %0, %8; addc.cc.u32 %1, %1, %9; addc.cc.u32 %2, %2, %10; addc.cc.u32 %3, %3, %11; " \
               "addc.cc.u32 %4, %4, %12; addc.cc.u32 %5, %5, %13; addc.cc.u32 %6, %6, %14; addc.u32 %7, %7, %15;"   \
               : "+r"(S[0]), "+r"(S[1]), "+r"(S[2]), "+r"(S[3]), "+r"(S[4]), "+r"(S[5]), "+r"(S[6]), "+r"(S[7])    \
               : "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]), "r"(x[4]), "r"(x[5]), "r"(x[6]), "r"(x[7]))
  if (WHICH == 0 || WHICH == 3 || WHICH == 6) {
    uint32_t lo[8], hi[8], s[8], t2[8];
#pragma unroll
    for (int j = 0; j < 8; j++) { lo[j] = x[j]; hi[j] = ~x[j]; s[j] = x[j] ^ 0x55u; t2[j] = x[j] ^ 0xaau; }
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int u = 0; u < 8; u++) {
#pragma unroll
        for (int j = 0; j < 8; j++)  // (hi:lo) += lo' * y : one IMAD.WIDE.U32, no carry chain
          asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;"
                       : "+r"(lo[j]), "+r"(hi[j]) : "r"(hi[(j + 3) & 7]), "r"(y));
        if (WHICH == 3 || WHICH == 6) ECB_ADD8(s);  // 8 x IADD3(.X)
        if (WHICH == 6) ECB_ADD8(t2);
      }
    }
    uint32_t t = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) t += lo[j] + hi[j] + s[j] + t2[j];
    out[tid] = t;
  } else if (WHICH == 1) {
    uint32_t lo[8], hi[8];
#pragma unroll
    for (int j = 0; j < 8; j++) { lo[j] = x[j]; hi[j] = ~x[j]; }
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int u = 0; u < 8; u++) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
          asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(lo[j]) : "r"(x[j]), "r"(y));  // IMAD
          asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(hi[j]) : "r"(x[j]), "r"(y));  // IMAD.HI.U32
        }
      }
    }
    uint32_t t = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) t += lo[j] ^ hi[j];
    out[tid] = t;
  } else if (WHICH == 2) {
    uint32_t s[8], t2[8];
#pragma unroll
    for (int j = 0; j < 8; j++) { s[j] = x[j]; t2[j] = ~x[j]; }
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int u = 0; u < 8; u++) {
        ECB_ADD8(s);
        ECB_ADD8(t2);
      }
    }
    uint32_t t = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) t ^= s[j] + t2[j];
    out[tid] = t;
  } else if (WHICH == 4) {
    // two independent 4-long IMAD.WIDE.U32.X carry chains per step (E/O accumulators)
    uint32_t e[8], o[8];
#pragma unroll
    for (int j = 0; j < 8; j++) { e[j] = x[j]; o[j] = ~x[j]; }
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int u = 0; u < 8; u++) {
        asm("mad.lo.cc.u32 %0, %8, %12, %0; madc.hi.cc.u32 %1, %8, %12, %1; madc.lo.cc.u32 %2, %9, %12, %2; madc.hi.cc.u32 %3, %9, %12, %3; "
            "madc.lo.cc.u32 %4, %10, %12, %4; madc.hi.cc.u32 %5, %10, %12, %5; madc.lo.cc.u32 %6, %11, %12, %6; madc.hi.u32 %7, %11, %12, %7;"
            : "+r"(e[0]), "+r"(e[1]), "+r"(e[2]), "+r"(e[3]), "+r"(e[4]), "+r"(e[5]), "+r"(e[6]), "+r"(e[7])
            : "r"(x[0]), "r"(x[2]), "r"(x[4]), "r"(x[6]), "r"(y));
        asm("mad.lo.cc.u32 %0, %8, %12, %0; madc.hi.cc.u32 %1, %8, %12, %1; madc.lo.cc.u32 %2, %9, %12, %2; madc.hi.cc.u32 %3, %9, %12, %3; "
            "madc.lo.cc.u32 %4, %10, %12, %4; madc.hi.cc.u32 %5, %10, %12, %5; madc.lo.cc.u32 %6, %11, %12, %6; madc.hi.u32 %7, %11, %12, %7;"
            : "+r"(o[0]), "+r"(o[1]), "+r"(o[2]), "+r"(o[3]), "+r"(o[4]), "+r"(o[5]), "+r"(o[6]), "+r"(o[7])
            : "r"(x[1]), "r"(x[3]), "r"(x[5]), "r"(x[7]), "r"(y));
      }
    }
    uint32_t t = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) t += e[j] ^ o[j];
    out[tid] = t;
  } else if (WHICH == 5) {
    int s[8];
#pragma unroll
    for (int j = 0; j < 8; j++) s[j] = (int)x[j];
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int u = 0; u < 16; u++) {
#pragma unroll
        for (int j = 0; j < 8; j++) s[j] = __vimax3_s32(s[j], (int)x[(j + 1) & 7], s[(j + 3) & 7]);
      }
    }
    int t = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) t ^= s[j];
    out[tid] = (uint32_t)t;
  }
}


// Generic pipe-mix probe: per step NW x IMAD.WIDE.U32, NL x IMAD (lo), NH x IMAD.HI.U32 and
// NA x (IADD3 + IADD3.X) pairs, every op on its own dependent chain (8 chains per kind).
template <int NW, int NL, int NH, int NA, int NM = 0>
__global__ void __launch_bounds__(256) k_mix(uint32_t* __restrict__ out, const uint32_t* __restrict__ in, int iters) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t wl[8], wh[8], l[8], h[8], al[8], ah[8], ml[8], mh[8];
  const uint32_t y = in[8];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    wl[j] = in[j] + tid; wh[j] = ~wl[j]; l[j] = wl[j] * 3u; h[j] = wl[j] * 5u; al[j] = wl[j] * 7u; ah[j] = wl[j] * 9u;
    ml[j] = wl[j] * 11u; mh[j] = wl[j] * 13u;
  }
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
#pragma unroll
      for (int j = 0; j < 12; j++) {
        if (j < NW) asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(wl[j & 7]), "+r"(wh[j & 7]) : "r"(wh[(j + 3) & 7]), "r"(y));
        if (j < NL) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(l[j & 7]) : "r"(y), "r"(l[(j + 3) & 7]));
        if (j < NH) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(h[j & 7]) : "r"(y), "r"(h[(j + 3) & 7]));
        if (j < NM) asm volatile("mul.lo.u32 %0, %2, %3; mul.hi.u32 %1, %2, %3;" : "=r"(ml[j & 7]), "=r"(mh[j & 7]) : "r"(mh[(j + 3) & 7] | 1u), "r"(ml[(j + 5) & 7]));
        if (j < NA) asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(al[j & 7]), "+r"(ah[j & 7]) : "r"(ah[(j + 3) & 7]), "r"(al[(j + 5) & 7]));
      }
    }
  }
  uint32_t t = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) t += wl[j] ^ wh[j] ^ l[j] ^ h[j] ^ al[j] ^ ah[j] ^ ml[j] ^ mh[j];
  out[tid] = t;
}

// Warp-specialised probe: even warps run only IMAD.WIDE chains, odd warps only IADD3 carry chains
// (MODE 0), or every warp runs both interleaved (MODE 1), same total work per pair of warps.
template <int MODE>
__global__ void __launch_bounds__(256) k_wspec(uint32_t* __restrict__ out, const uint32_t* __restrict__ in, int iters) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  const bool wwarp = ((threadIdx.x >> 7) & 1) == 0;  // warps 0-3 multiply, 4-7 add: one of each kind per sub-partition pair
  uint32_t wl[8], wh[8], al[8], ah[8];
  const uint32_t y = in[8];
#pragma unroll
  for (int j = 0; j < 8; j++) { wl[j] = in[j] + tid; wh[j] = ~wl[j]; al[j] = wl[j] * 7u; ah[j] = wl[j] * 9u; }
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      if (MODE == 1 || wwarp) {
#pragma unroll
        for (int j = 0; j < (MODE == 1 ? 8 : 16); j++)
          asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(wl[j & 7]), "+r"(wh[j & 7]) : "r"(wh[(j + 3) & 7]), "r"(y));
      }
      if (MODE == 1 || !wwarp) {
#pragma unroll
        for (int j = 0; j < (MODE == 1 ? 12 : 24); j++)
          asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(al[j & 7]), "+r"(ah[j & 7]) : "r"(ah[(j + 3) & 7]), "r"(al[(j + 5) & 7]));
      }
    }
  }
  uint32_t t = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) t += wl[j] ^ wh[j] ^ al[j] ^ ah[j];
  out[tid] = t;
}

struct MixCombo { int nw, nl, nh, na; void (*fn)(uint32_t*, const uint32_t*, int); };
#define MIX(a, b, c, d) {a, b, c, d, k_mix<a, b, c, d>}
static const MixCombo g_mix[] = {
    MIX(8, 0, 0, 0), MIX(0, 8, 0, 0), MIX(0, 0, 8, 0), MIX(0, 8, 8, 0), MIX(0, 0, 0, 8), MIX(8, 0, 0, 4), MIX(8, 0, 0, 8),
    MIX(8, 0, 0, 12), MIX(0, 8, 8, 8), MIX(0, 8, 0, 8), MIX(0, 0, 8, 8), MIX(8, 8, 0, 0), MIX(8, 0, 8, 0), MIX(4, 8, 8, 8),
    MIX(8, 4, 0, 8), MIX(8, 0, 4, 8), MIX(4, 0, 0, 12), MIX(0, 12, 0, 0), MIX(0, 12, 0, 12),
    {-1, 0, 0, 0, k_wspec<0>}, {-2, 0, 0, 0, k_wspec<1>},
    {0, 0, 0, 0, k_mix<0, 0, 0, 0, 8>}, {0, 0, 0, 8, k_mix<0, 0, 0, 8, 8>}, {0, 0, 0, 12, k_mix<0, 0, 0, 12, 8>}, {4, 0, 0, 8, k_mix<4, 0, 0, 8, 4>}};
static const int g_nmix = (int)(sizeof(g_mix) / sizeof(g_mix[0]));

template <int L, int OP, bool Q>
static int launch_field(void* out, const void* a, const void* b, size_t n, int param, cudaStream_t s) {
  if (n == 0) return ECB200_OK;
  const int threads = 256;
  const unsigned blocks = (unsigned)((n + threads - 1) / threads);
  k_field<L, OP, Q><<<blocks, threads, 0, s>>>(out, a, b, n, param);
  ECB_LAUNCH_CHECK();
  return ECB200_OK;
}
template <int OP, bool Q>
static int launch_field_l(int L, void* out, const void* a, const void* b, size_t n, int param, cudaStream_t s) {
  switch (L) {
    case L_LANE: return launch_field<L_LANE, OP, Q>(out, a, b, n, param, s);
    case L_PACK4: return launch_field<L_PACK4, OP, Q>(out, a, b, n, param, s);
    default: return launch_field<L_SOA, OP, Q>(out, a, b, n, param, s);
  }
}
template <int OP>
static int launch_field_q(bool q, int L, void* out, const void* a, const void* b, size_t n, int param, cudaStream_t s) {
  // only squaring-based ops have a quirk variant
  if (OP == OP_SQR || OP == OP_INV) {
    return q ? launch_field_l<OP, true>(L, out, a, b, n, param, s) : launch_field_l<OP, false>(L, out, a, b, n, param, s);
  }
  return launch_field_l<OP, true>(L, out, a, b, n, param, s);
}

// Field ops work directly on the caller's layout (no conversion pass); host memory is
// staged through stream-ordered device temporaries.
template <int OP>
static int field_call(void* out, const void* a, const void* b, size_t n, int param, uint32_t flags, void* stream) {
  int rc = check_common(n, flags);
  if (rc) return rc;
  if (n == 0) return ECB200_OK;
  if (!out || !a || ((OP == OP_ADD || OP == OP_SUB || OP == OP_MUL || OP == OP_MULCHAIN) && !b)) {
    set_error("null pointer argument");
    return ECB200_ERR_ARG;
  }
  cudaStream_t s = (cudaStream_t)stream;
  const int L = layout_of(flags);
  const size_t bytes = operand_bytes(n, 1);
  if (on_device(flags)) return launch_field_q<OP>(quirk_on(flags), L, out, a, b, n, param, s);
  Scratch sc(s);
  void *da = nullptr, *db = nullptr, *dout = nullptr;
  if ((rc = sc.alloc(&da, bytes))) return rc;
  if ((rc = sc.alloc(&dout, bytes))) return rc;
  ECB_CUDA(cudaMemcpyAsync(da, a, bytes, cudaMemcpyHostToDevice, s));
  if (b) {
    if ((rc = sc.alloc(&db, bytes))) return rc;
    ECB_CUDA(cudaMemcpyAsync(db, b, bytes, cudaMemcpyHostToDevice, s));
  }
  if ((rc = launch_field_q<OP>(quirk_on(flags), L, dout, da, db, n, param, s))) return rc;
  ECB_CUDA(cudaMemcpyAsync(out, dout, bytes, cudaMemcpyDeviceToHost, s));
  ECB_CUDA(cudaStreamSynchronize(s));
  return ECB200_OK;
}

// layout conversion helpers used by the point/scalar TU
int convert_to_soa(int L, void* dst, const void* src, size_t n, int nc, cudaStream_t s) {
  if (n == 0) return ECB200_OK;
  const unsigned blocks = (unsigned)((n + 255) / 256);
  if (L == L_LANE) k_to_soa<L_LANE><<<blocks, 256, 0, s>>>(dst, src, n, nc);
  else if (L == L_PACK4) k_to_soa<L_PACK4><<<blocks, 256, 0, s>>>(dst, src, n, nc);
  else { set_error("convert_to_soa: bad layout"); return ECB200_ERR_ARG; }
  ECB_LAUNCH_CHECK();
  return ECB200_OK;
}
int convert_from_soa(int L, void* dst, const void* src, size_t n, int nc, cudaStream_t s) {
  if (n == 0) return ECB200_OK;
  const unsigned blocks = (unsigned)((n + 255) / 256);
  if (L == L_LANE) k_from_soa<L_LANE><<<blocks, 256, 0, s>>>(dst, src, n, nc);
  else if (L == L_PACK4) k_from_soa<L_PACK4><<<blocks, 256, 0, s>>>(dst, src, n, nc);
  else { set_error("convert_from_soa: bad layout"); return ECB200_ERR_ARG; }
  ECB_LAUNCH_CHECK();
  return ECB200_OK;
}

}  // namespace ecb200

using namespace ecb200;

template <int LS>
static int convert_dispatch(int LD, void* dst, const void* src, size_t n, int nc, cudaStream_t s) {
  const unsigned blocks = (unsigned)((n + 255) / 256);
  if (LD == L_LANE) k_convert<LS, L_LANE><<<blocks, 256, 0, s>>>(dst, src, n, nc);
  else if (LD == L_PACK4) k_convert<LS, L_PACK4><<<blocks, 256, 0, s>>>(dst, src, n, nc);
  else k_convert<LS, L_SOA><<<blocks, 256, 0, s>>>(dst, src, n, nc);
  ECB_LAUNCH_CHECK();
  return ECB200_OK;
}

template <bool TO_BYTES>
static int bytes_call(void* vals, void* bytes, int ncoord, size_t n, uint32_t flags, void* stream) {
  int rc = check_common(n, flags);
  if (rc) return rc;
  if (ncoord < 1 || ncoord > 3 || !vals || !bytes) { set_error("bad argument to ecb200_bn_*_bytes_be"); return ECB200_ERR_ARG; }
  if (n == 0) return ECB200_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const size_t nb = operand_bytes(n, ncoord);
  Scratch sc(s);
  void *dv = vals, *db = bytes;
  if (!on_device(flags)) {
    if ((rc = sc.alloc(&dv, nb)) || (rc = sc.alloc(&db, nb))) return rc;
    ECB_CUDA(cudaMemcpyAsync(TO_BYTES ? dv : db, TO_BYTES ? vals : bytes, nb, cudaMemcpyHostToDevice, s));
  }
  const unsigned blocks = (unsigned)((n + 255) / 256);
  const int L = layout_of(flags);
  if (L == L_LANE) k_bytes_be<L_LANE, TO_BYTES><<<blocks, 256, 0, s>>>(dv, (uint8_t*)db, n, ncoord);
  else if (L == L_PACK4) k_bytes_be<L_PACK4, TO_BYTES><<<blocks, 256, 0, s>>>(dv, (uint8_t*)db, n, ncoord);
  else k_bytes_be<L_SOA, TO_BYTES><<<blocks, 256, 0, s>>>(dv, (uint8_t*)db, n, ncoord);
  ECB_LAUNCH_CHECK();
  if (!on_device(flags)) {
    ECB_CUDA(cudaMemcpyAsync(TO_BYTES ? bytes : vals, TO_BYTES ? db : dv, nb, cudaMemcpyDeviceToHost, s));
    ECB_CUDA(cudaStreamSynchronize(s));
  }
  return ECB200_OK;
}
namespace ecb200 {
int release_device_resources(DeviceCtx* c);   // kernels_point.cu: fixed-base tables, pipeline streams, bounce buffers
int prebuild_base_tables(DeviceCtx* c);
}

extern "C" {

int ecb200_abi_version(void) { return ECB200_ABI_VERSION; }
const char* ecb200_last_error(void) { return t_err; }
uint64_t ecb200_launch_count(void) { return g_launches.load(); }

int ecb200_init(int device) {
  int count = 0;
  ECB_CUDA(cudaGetDeviceCount(&count));
  if (device < 0 || device >= count) {
    set_error("device %d out of range (%d CUDA devices)", device, count);
    return ECB200_ERR_ARG;
  }
  ECB_CUDA(cudaSetDevice(device));
  DeviceCtx* c = nullptr;
  int rc = ensure_ctx(device, &c);
  if (rc) return rc;
  // the fixed-base tables are built here, so that no later ECB200_MEM_DEVICE call has to synchronise
  return prebuild_base_tables(c);
}

int ecb200_init_devices(const int* devices, int count) {
  if (count < 0 || (count > 0 && !devices)) {
    set_error("bad device list");
    return ECB200_ERR_ARG;
  }
  for (int i = 0; i < count; i++)
    for (int j = 0; j < i; j++)
      if (devices[i] == devices[j]) {
        set_error("device %d listed twice", devices[i]);
        return ECB200_ERR_ARG;
      }
  for (int i = count - 1; i >= 0; i--) {   // ends on devices[0]: it becomes the calling thread's device
    int rc = ecb200_init(devices[i]);
    if (rc) return rc;
  }
  std::lock_guard<std::mutex> lock(g_multi_mu);
  g_multi.assign(devices, devices + count);
  if (count <= 1) g_multi.clear();
  return ECB200_OK;
}

int ecb200_device_count(void) {
  std::lock_guard<std::mutex> lock(g_multi_mu);
  return g_multi.empty() ? 1 : (int)g_multi.size();
}

int ecb200_shutdown(void) {
  int device = 0;
  ECB_CUDA(cudaGetDevice(&device));
  ECB_CUDA(cudaDeviceSynchronize());
  if (device < 0 || device >= kMaxDevices) return ECB200_OK;
  {
    std::lock_guard<std::mutex> lock(g_multi_mu);
    for (size_t i = 0; i < g_multi.size(); i++)
      if (g_multi[i] == device) { g_multi.clear(); break; }   // back to single-device dispatch
  }
  DeviceCtx& c = g_ctx[device];
  int rc = release_device_resources(&c);
  if (rc) return rc;
  std::lock_guard<std::mutex> lock(c.mu);
  if (c.ready) {
    ECB_CUDA(cudaMemPoolDestroy(c.pool));
    c.pool = nullptr;
    c.ready = false;   // re-created by the next call on this device
  }
  return ECB200_OK;
}

int ecb200_mgry_add(void* out, const void* a, const void* b, size_t n, uint32_t flags, void* stream) {
  return field_call<OP_ADD>(out, a, b, n, 0, flags, stream);
}
int ecb200_mgry_sub(void* out, const void* a, const void* b, size_t n, uint32_t flags, void* stream) {
  return field_call<OP_SUB>(out, a, b, n, 0, flags, stream);
}
int ecb200_mgry_mul(void* out, const void* a, const void* b, size_t n, uint32_t flags, void* stream) {
  return field_call<OP_MUL>(out, a, b, n, 0, flags, stream);
}
int ecb200_mgry_sqr(void* out, const void* a, size_t n, uint32_t flags, void* stream) {
  return field_call<OP_SQR>(out, a, nullptr, n, 0, flags, stream);
}
int ecb200_mgry_shift_left(void* out, const void* a, int count, size_t n, uint32_t flags, void* stream) {
  if (count < 1 || count > 8) {
    set_error("shift count %d out of range 1..8", count);
    return ECB200_ERR_ARG;
  }
  return field_call<OP_SHL>(out, a, nullptr, n, count, flags, stream);
}
int ecb200_gfp_opposite(void* out, const void* a, size_t n, uint32_t flags, void* stream) {
  return field_call<OP_NEG>(out, a, nullptr, n, 0, flags, stream);
}
int ecb200_from_classical(void* out, const void* a, size_t n, uint32_t flags, void* stream) {
  return field_call<OP_FROMC>(out, a, nullptr, n, 0, flags, stream);
}
int ecb200_to_classical(void* out, const void* a, size_t n, uint32_t flags, void* stream) {
  return field_call<OP_TOC>(out, a, nullptr, n, 0, flags, stream);
}
int ecb200_gfp_inverse(void* out, const void* a, size_t n, uint32_t flags, void* stream) {
  return field_call<OP_INV>(out, a, nullptr, n, 0, flags, stream);
}
int ecb200_mgry_mul_chain(void* out, const void* a, const void* b, int iters, size_t n, uint32_t flags, void* stream) {
  if (iters < 0) {
    set_error("iters must be >= 0");
    return ECB200_ERR_ARG;
  }
  return field_call<OP_MULCHAIN>(out, a, b, n, iters, flags, stream);
}

int ecb200_synth_values(void* out, uint64_t seed, uint64_t start, int kind, size_t n, uint32_t flags, void* stream) {
  int rc = check_common(n, flags);
  if (rc) return rc;
  if (n == 0) return ECB200_OK;
  if (!out || (kind != 0 && kind != 1)) {
    set_error("bad argument to ecb200_synth_values");
    return ECB200_ERR_ARG;
  }
  cudaStream_t s = (cudaStream_t)stream;
  const int L = layout_of(flags);
  const unsigned blocks = (unsigned)((n + 255) / 256);
  Scratch sc(s);
  void* d = out;
  if (!on_device(flags) && (rc = sc.alloc(&d, operand_bytes(n, 1)))) return rc;
  if (L == L_LANE) k_synth<L_LANE><<<blocks, 256, 0, s>>>(d, seed, start, kind, n);
  else if (L == L_PACK4) k_synth<L_PACK4><<<blocks, 256, 0, s>>>(d, seed, start, kind, n);
  else k_synth<L_SOA><<<blocks, 256, 0, s>>>(d, seed, start, kind, n);
  ECB_LAUNCH_CHECK();
  if (!on_device(flags)) {
    ECB_CUDA(cudaMemcpyAsync(out, d, operand_bytes(n, 1), cudaMemcpyDeviceToHost, s));
    ECB_CUDA(cudaStreamSynchronize(s));
  }
  return ECB200_OK;
}

int ecb200_checksum(uint32_t* out8, const void* buf, size_t nwords, void* stream) {
  if (!out8 || (!buf && nwords) || (nwords & 7)) {
    set_error("ecb200_checksum: need a device buffer whose word count is a multiple of 8");
    return ECB200_ERR_ARG;
  }
  cudaStream_t s = (cudaStream_t)stream;
  Scratch sc(s);
  void* acc = nullptr;
  int rc = sc.alloc(&acc, 32);
  if (rc) return rc;
  ECB_CUDA(cudaMemsetAsync(acc, 0, 32, s));
  if (nwords) {
    size_t groups = nwords / 8;
    unsigned blocks = (unsigned)((groups + 255) / 256);
    if (blocks > 148u * 16u) blocks = 148u * 16u;
    k_checksum<<<blocks, 256, 0, s>>>((uint32_t*)acc, (const uint32_t*)buf, nwords);
    ECB_LAUNCH_CHECK();
  }
  ECB_CUDA(cudaMemcpyAsync(out8, acc, 32, cudaMemcpyDeviceToHost, s));
  ECB_CUDA(cudaStreamSynchronize(s));
  return ECB200_OK;
}

int ecb200_microbench(int which, int blocks, int threads, int iters, double* ops_per_iter, float* ms, void* stream) {
  if (which < 0 || which > 6 || blocks < 1 || threads < 32 || threads > 256 || iters < 1 || !ops_per_iter || !ms) {
    set_error("bad argument to ecb200_microbench");
    return ECB200_ERR_ARG;
  }
  cudaStream_t s = (cudaStream_t)stream;
  Scratch sc(s);
  void *din = nullptr, *dout = nullptr;
  int rc;
  if ((rc = sc.alloc(&din, 64))) return rc;
  if ((rc = sc.alloc(&dout, (size_t)blocks * threads * 4))) return rc;
  const uint32_t seedv[9] = {0x9e3779b9u, 0x7f4a7c15u, 0xf39cc060u, 0x5cedc834u, 0x1082276bu, 0xf3a27251u, 0xf86c6a11u, 0xd0c18e95u, 0x2767f0b1u};
  ECB_CUDA(cudaMemcpyAsync(din, seedv, sizeof seedv, cudaMemcpyHostToDevice, s));
  cudaEvent_t e0, e1;
  ECB_CUDA(cudaEventCreate(&e0));
  ECB_CUDA(cudaEventCreate(&e1));
  for (int rep = 0; rep < 2; rep++) {  // first pass warms up
    ECB_CUDA(cudaEventRecord(e0, s));
    switch (which) {
      case 0: k_microbench<0><<<blocks, threads, 0, s>>>((uint32_t*)dout, (const uint32_t*)din, iters); break;
      case 1: k_microbench<1><<<blocks, threads, 0, s>>>((uint32_t*)dout, (const uint32_t*)din, iters); break;
      case 2: k_microbench<2><<<blocks, threads, 0, s>>>((uint32_t*)dout, (const uint32_t*)din, iters); break;
      case 3: k_microbench<3><<<blocks, threads, 0, s>>>((uint32_t*)dout, (const uint32_t*)din, iters); break;
      case 4: k_microbench<4><<<blocks, threads, 0, s>>>((uint32_t*)dout, (const uint32_t*)din, iters); break;
      case 5: k_microbench<5><<<blocks, threads, 0, s>>>((uint32_t*)dout, (const uint32_t*)din, iters); break;
      default: k_microbench<6><<<blocks, threads, 0, s>>>((uint32_t*)dout, (const uint32_t*)din, iters); break;
    }
    ECB_LAUNCH_CHECK();
    ECB_CUDA(cudaEventRecord(e1, s));
    ECB_CUDA(cudaEventSynchronize(e1));
  }
  ECB_CUDA(cudaEventElapsedTime(ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  // counted instructions per thread per loop trip
  const double per[7] = {64, 128, 120, 124, 64, 128, 184};  // counted in the SASS of each loop (tools/ + cuobjdump)
  *ops_per_iter = per[which];
  return ECB200_OK;
}

int ecb200_convert_layout(void* dst, uint32_t dst_layout, const void* src, uint32_t src_layout, int ncoord, size_t n, uint32_t flags, void* stream) {
  int rc = check_common(n, src_layout | (flags & ~ECB200_LAYOUT_MASK));
  if (!rc) rc = check_common(n, dst_layout | (flags & ~ECB200_LAYOUT_MASK));
  if (rc) return rc;
  if (ncoord < 1 || ncoord > 3 || !dst || !src) { set_error("bad argument to ecb200_convert_layout"); return ECB200_ERR_ARG; }
  if (n == 0) return ECB200_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const size_t bytes = operand_bytes(n, ncoord);
  Scratch sc(s);
  const void* dsrc = src;
  void* ddst = dst;
  if (!on_device(flags)) {
    void *a, *b;
    if ((rc = sc.alloc(&a, bytes)) || (rc = sc.alloc(&b, bytes))) return rc;
    ECB_CUDA(cudaMemcpyAsync(a, src, bytes, cudaMemcpyHostToDevice, s));
    dsrc = a; ddst = b;
  }
  const int LS = (int)(src_layout & ECB200_LAYOUT_MASK), LD = (int)(dst_layout & ECB200_LAYOUT_MASK);
  if (LS == L_LANE) rc = convert_dispatch<L_LANE>(LD, ddst, dsrc, n, ncoord, s);
  else if (LS == L_PACK4) rc = convert_dispatch<L_PACK4>(LD, ddst, dsrc, n, ncoord, s);
  else rc = convert_dispatch<L_SOA>(LD, ddst, dsrc, n, ncoord, s);
  if (rc) return rc;
  if (!on_device(flags)) {
    ECB_CUDA(cudaMemcpyAsync(dst, ddst, bytes, cudaMemcpyDeviceToHost, s));
    ECB_CUDA(cudaStreamSynchronize(s));
  }
  return ECB200_OK;
}

int ecb200_bn_from_bytes_be(void* vals, const void* bytes, int ncoord, size_t n, uint32_t flags, void* stream) {
  return bytes_call<false>(vals, const_cast<void*>(bytes), ncoord, n, flags, stream);
}
int ecb200_bn_to_bytes_be(void* bytes, const void* vals, int ncoord, size_t n, uint32_t flags, void* stream) {
  return bytes_call<true>(const_cast<void*>(vals), bytes, ncoord, n, flags, stream);
}

int ecb200_microbench_mix(int combo, int blocks, int threads, int iters, int* counts4, float* ms, void* stream) {
  if (combo < 0 || combo >= g_nmix) { set_error("combo out of range (0..%d)", g_nmix - 1); return ECB200_ERR_ARG; }
  if (blocks < 1 || threads < 32 || threads > 256 || iters < 1 || !counts4 || !ms) { set_error("bad argument"); return ECB200_ERR_ARG; }
  cudaStream_t s = (cudaStream_t)stream;
  Scratch sc(s);
  void *din = nullptr, *dout = nullptr;
  int rc;
  if ((rc = sc.alloc(&din, 64))) return rc;
  if ((rc = sc.alloc(&dout, (size_t)blocks * threads * 4))) return rc;
  const uint32_t seedv[9] = {0x9e3779b9u, 0x7f4a7c15u, 0xf39cc060u, 0x5cedc834u, 0x1082276bu, 0xf3a27251u, 0xf86c6a11u, 0xd0c18e95u, 0x2767f0b1u};
  ECB_CUDA(cudaMemcpyAsync(din, seedv, sizeof seedv, cudaMemcpyHostToDevice, s));
  cudaEvent_t e0, e1;
  ECB_CUDA(cudaEventCreate(&e0));
  ECB_CUDA(cudaEventCreate(&e1));
  for (int rep = 0; rep < 2; rep++) {
    ECB_CUDA(cudaEventRecord(e0, s));
    g_mix[combo].fn<<<blocks, threads, 0, s>>>((uint32_t*)dout, (const uint32_t*)din, iters);
    ECB_LAUNCH_CHECK();
    ECB_CUDA(cudaEventRecord(e1, s));
    ECB_CUDA(cudaEventSynchronize(e1));
  }
  ECB_CUDA(cudaEventElapsedTime(ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  counts4[0] = g_mix[combo].nw; counts4[1] = g_mix[combo].nl; counts4[2] = g_mix[combo].nh; counts4[3] = g_mix[combo].na;
  return ECB200_OK;
}
int ecb200_microbench_mix_count(void) { return g_nmix; }

}  // extern "C"
