// kernels_point.cu -- co-Z point kernels and the batched scalar multiplication, one
// lane per thread, plus their C-ABI entry points (include/ecb200.h).
//
// All point kernels read and write the device-native planar layout (ECB200_LAYOUT_SOA:
// coalesced 128-bit accesses); other layouts and host memory go through the conversion
// kernels of kernels_field.cu.  The scalar multiplication is integer-multiply bound
// (211 540 MAC32 per lane against 224 bytes of traffic), so the conversion passes are
// noise there.
#include <cstdlib>
#include <mutex>
#include <string>
#include <thread>

#include "host_common.cuh"
#include "layout.cuh"
#include "point.cuh"

namespace ecb200 {

int convert_to_soa(int L, void* dst, const void* src, size_t n, int nc, cudaStream_t s);
int convert_from_soa(int L, void* dst, const void* src, size_t n, int nc, cudaStream_t s);

using S = Layout<L_SOA>;

__device__ __forceinline__ jac load_jac(const void* p, size_t n, size_t i) {
  jac r;
  r.x = S::load(p, n, i, 3, 0);
  r.y = S::load(p, n, i, 3, 1);
  r.z = S::load(p, n, i, 3, 2);
  return r;
}
__device__ __forceinline__ void store_jac(void* p, size_t n, size_t i, const jac& a) {
  S::store(p, n, i, 3, 0, a.x);
  S::store(p, n, i, 3, 1, a.y);
  S::store(p, n, i, 3, 2, a.z);
}

enum PointOp : int { PO_DBLU, PO_ZADDU, PO_ZDAU, PO_ADDZ21, PO_TRPLU };

// One point op per thread.  The op runs in Lazy mode (no branches on the 2^-32 cases); a lane that
// flagged one is re-run from its (reloaded) inputs in Exact mode by an out-of-line copy.
template <int OP, bool QUIRK, class MD>
__device__ __forceinline__ void point_compute(const void* A, const void* B, size_t n, size_t i, MD& md, jac& r, jac& w) {
  jac a = load_jac(A, n, i);
  // r: returned point, w: rewritten operand
  if (OP == PO_DBLU) { r = pt_dblu<QUIRK>(a, md); w = a; }
  else if (OP == PO_TRPLU) { r = pt_trplu<QUIRK>(a, md); w = a; }
  else if (OP == PO_ZADDU) { const jac b = load_jac(B, n, i); r = pt_zaddu<QUIRK>(a, b, md); w = a; }
  else if (OP == PO_ZDAU) { jac q = load_jac(B, n, i); r = pt_zdau<QUIRK>(a, q, md); w = q; }
  else { const fe bx = S::load(B, n, i, 3, 0), by = S::load(B, n, i, 3, 1); r = pt_add_z2_1<QUIRK>(a, bx, by, md); w = r; }
}
template <int OP>
__device__ __forceinline__ void point_store(void* out1, void* out2, size_t n, size_t i, const jac& r, const jac& w) {
  if (OP == PO_ADDZ21) store_jac(out1, n, i, r);
  else { store_jac(out1, n, i, w); store_jac(out2, n, i, r); }
}
template <int OP, bool QUIRK>
__device__ __noinline__ void point_op_exact(void* out1, void* out2, const void* A, const void* B, size_t n, size_t i) {
  Exact md;
  jac r, w;
  point_compute<OP, QUIRK>(A, B, n, i, md, r, w);
  point_store<OP>(out1, out2, n, i, r, w);
}

// Block size per op: the multiply-bound ZDAU (50 KB of straight-line code) runs as two 256-thread blocks
// per SM: warps of a block stay close together and share instruction-cache lines, and while one block
// loads its 384 bytes per point or stores its results the other one multiplies (one 512-thread block per SM
// serialised load -> compute -> store: 0.47 of the IMAD peak against the ladder's 0.57 for the same formula).
// The shorter ops, which lean on HBM, prefer many small blocks.
#ifndef ECB200_ZDAU_THREADS
#define ECB200_ZDAU_THREADS 256
#define ECB200_ZDAU_BLOCKS 2
#endif
template <int OP>
struct PointThreads { static constexpr int value = (OP == PO_ZDAU) ? ECB200_ZDAU_THREADS : 128; static constexpr int blocks = (OP == PO_ZDAU) ? ECB200_ZDAU_BLOCKS : 1; };
template <int OP, bool QUIRK>
__global__ void __launch_bounds__(PointThreads<OP>::value, PointThreads<OP>::blocks) k_point(void* out1, void* out2, const void* A, const void* B, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Lazy md;
  jac r, w;
  point_compute<OP, QUIRK>(A, B, n, i, md, r, w);
  // nothing has been stored yet, so the exact re-run still sees the inputs even if outputs alias them
  if (__builtin_expect(md.flagged(), 0)) point_op_exact<OP, QUIRK>(out1, out2, A, B, n, i);
  else point_store<OP>(out1, out2, n, i, r, w);
}

// The ladder kernel.  mode 0: per-lane point P[i]; mode 1: P = G for every lane (with TABW > 0 the
// ladder starts from the table of states after TABW bits); the scalar is per lane unless k_bcast
// (scalar_mult_1s: one scalar for all lanes).  One block of 512 threads per SM (16 warps, 128
// registers each), all warps kept in step by a barrier per ladder iteration: the ~50 KB loop body
// is then fetched once per SM instead of once per warp.
// L: layout of k, P and out -- the ladder reads 128 and writes 96 bytes per lane, so it takes the
// caller's layout directly (no conversion pass, no extra kernels competing for the SMs).
// Scalar words and P are re-read from memory where they are needed (SrcGlobal) rather than held
// in registers through the loop.
#ifndef SK_STRIDE
#define SK_STRIDE blockDim.x
#endif
// The scalar is staged once in shared memory (8 x THREADS words, each thread its own column), so that the
// loop's only scalar input is one shared-memory address: the same loop for every layout (the pointer arithmetic of the
// pack layout cost ptxas 50 instructions more per step than the other two).
template <int MODE, int L, int THREADS>
struct SrcGlobal {
  const void* k;
  const void* P;
  const uint4* tab;
  size_t n, i;
  int k_bcast;
  const uint32_t* sk;   // shared: word w of this lane's scalar at sk[w * THREADS], w = 0..7
  __device__ __forceinline__ uint32_t kword_global(int w) const {
    return k_bcast ? Layout<L_LANE>::load_word(k, 1, 0, 1, 0, w) : Layout<L>::load_word(k, n, i, 1, 0, w);
  }
  __device__ __forceinline__ uint32_t kword(int w) const { return sk[w * SK_STRIDE]; }
  // the lane index, recomputed from the special registers and made opaque: nothing derived from it (the pack
  // layout's (i >> 2, i & 3) pair, an SOA row address) is kept in registers through the loop for the code after it
  static __device__ __forceinline__ size_t lane_index(size_t n) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    asm volatile("" : "+l"(t));
    return t < n ? t : n - 1;
  }
  __device__ __forceinline__ void point(fe& x, fe& y) const {
    if (MODE == 0) {
      const void* p = P;
      asm volatile("" : "+l"(p));  // a fresh load each time: the value is not kept live through the loop
      const size_t ii = lane_index(n);
      x = Layout<L>::load(p, n, ii, 3, 0);
      y = Layout<L>::load(p, n, ii, 3, 1);
    } else {
      const uint32_t gx[8] = ECB200_GXM_WORDS, gy[8] = ECB200_GYM_WORDS;
      x = fe_const(gx);
      y = fe_const(gy);
    }
  }
  __device__ __forceinline__ void table(uint32_t idx, fe (&st)[5]) const {
    const uint4* e = tab + (size_t)idx * 10;
#pragma unroll
    for (int c = 0; c < 5; c++) {
      const uint4 lo = __ldg(e + 2 * c), hi = __ldg(e + 2 * c + 1);
      st[c].v[0] = lo.x; st[c].v[1] = lo.y; st[c].v[2] = lo.z; st[c].v[3] = lo.w;
      st[c].v[4] = hi.x; st[c].v[5] = hi.y; st[c].v[6] = hi.z; st[c].v[7] = hi.w;
    }
  }
};

template <bool QUIRK, int MODE, int THREADS, int L, int TABW>
__global__ void __launch_bounds__(THREADS, 1) k_scalar_mult_sync(void* __restrict__ out, const void* __restrict__ k,
                                                                 const void* __restrict__ P, size_t n, int k_bcast,
                                                                 const uint4* __restrict__ tab) {
  const size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t i = i0 < n ? i0 : n - 1;  // surplus threads recompute the last lane (they must reach the barriers)
  __shared__ uint32_t s_k[8 * THREADS];
  const SrcGlobal<MODE, L, THREADS> src{k, P, tab, n, i, k_bcast, s_k + threadIdx.x};
#pragma unroll
  for (int w = 0; w < 8; w++) s_k[w * THREADS + threadIdx.x] = src.kword_global(w);   // read back by the same thread only
  const jac r = pt_scalar_mult<QUIRK, true, TABW>(src);
  size_t j0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // recomputed, not carried through the loop
  asm volatile("" : "+l"(j0));
  if (j0 < n) {
    Layout<L>::store(out, n, j0, 3, 0, r.x);
    Layout<L>::store(out, n, j0, 3, 1, r.y);
    Layout<L>::store(out, n, j0, 3, 2, r.z);
  }
}

// Fixed-base table: entry idx (TABW bits) = ladder state of G after the steps for scalar bits
// 1..TABW = idx, five field elements (base.x, base.y, P.x, P.y, Z) of 32 bytes each.
template <bool QUIRK>
__global__ void __launch_bounds__(128) k_build_base_table(uint4* __restrict__ tab, int tabw) {
  const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (1u << tabw)) return;
  const uint32_t gx[8] = ECB200_GXM_WORDS, gy[8] = ECB200_GYM_WORDS;
  uint32_t xy[16], st[40];
  for (int j = 0; j < 8; j++) { xy[j] = gx[j]; xy[8 + j] = gy[j]; }
  pt_ladder_prefix_exact<QUIRK>(st, idx, tabw, xy);
  uint4* e = tab + (size_t)idx * 10;
  for (int c = 0; c < 10; c++) e[c] = make_uint4(st[4 * c], st[4 * c + 1], st[4 * c + 2], st[4 * c + 3]);
}

// to_affine (jacobian_curve_point.h:33-42): invZ = z^(p-2); x = X*invZ^2; y = Y*invZ^3; to_classical.  255 squarings +
// 128 + 6 multiplications per lane = 17 792 algorithmic MAC32 against 160 bytes: integer-multiply bound.  Lazy mode
// with an out-of-line exact re-run of flagged lanes, like the point kernels.
template <bool QUIRK, class MD>
__device__ __forceinline__ void to_affine_compute(const jac& a, fe& x, fe& y, MD& md) {
  const fe iz = fp_inv<QUIRK>(a.z, md);
  const fe iz2 = fp_sqr<QUIRK>(iz, md);
  const fe iz3 = fp_mul(iz2, iz, md);
  x = fp_to_classical(fp_mul(a.x, iz2, md), md);
  y = fp_to_classical(fp_mul(a.y, iz3, md), md);
}
template <bool QUIRK>
__device__ __noinline__ void to_affine_exact(void* xy, const void* J, size_t n, size_t i) {
  Exact md;
  fe x, y;
  to_affine_compute<QUIRK>(load_jac(J, n, i), x, y, md);
  S::store(xy, n, i, 2, 0, x);
  S::store(xy, n, i, 2, 1, y);
}
template <bool QUIRK>
__global__ void __launch_bounds__(256, 2) k_to_affine(void* __restrict__ xy, const void* __restrict__ J, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Lazy md;
  fe x, y;
  to_affine_compute<QUIRK>(load_jac(J, n, i), x, y, md);
  if (__builtin_expect(md.flagged(), 0)) { to_affine_exact<QUIRK>(xy, J, n, i); return; }
  S::store(xy, n, i, 2, 0, x);
  S::store(xy, n, i, 2, 1, y);
}

// wide_curve_point::from_x / curve_group::compute_y (curve_point_ops.h:12-22, curve_group.h:43-58):
// y = sqrt(x^3 - 3x + b) with sqrt = pow((p+1)/4) (gfp.h:46-54) through the reference's LSB-first
// square-and-multiply (mgry_ops.h:44-86, same sequence of squarings), then the check r^2 == y^2.
// ok[i] = 1 iff lane i has a square root (the reference answers per 4-lane pack: all four or none).
// 255 squarings + 6 multiplications = 9 564 algorithmic MAC32 per lane.
template <bool QUIRK, class MD>
__device__ __forceinline__ void from_x_compute(const fe& xc, fe& y, uint8_t& ok, MD& md) {
  const fe xm = fp_from_classical(xc, md);
  const fe xpow3 = fp_mul(fp_sqr<QUIRK>(xm, md), xm, md);
  const fe x3 = fp_add(fp_shl1(xm, md), xm, md);
  const fe ypow2 = fp_sub(fp_add(xpow3, fe_BM(), md), x3);
  const uint32_t e[8] = ECB200_SQRT_EXP_WORDS;
  const fe res = fp_pow_lsb<QUIRK>(ypow2, e, 254, md);
  ok = fe_eq(fp_sqr<QUIRK>(res, md), ypow2) ? 1 : 0;
  y = fp_to_classical(res, md);
}
template <bool QUIRK>
__device__ __noinline__ void from_x_exact(void* y, uint8_t* ok, const void* x, size_t n, size_t i) {
  Exact md;
  fe r;
  uint8_t o;
  from_x_compute<QUIRK>(S::load(x, n, i, 1, 0), r, o, md);
  ok[i] = o;
  S::store(y, n, i, 1, 0, r);
}
template <bool QUIRK>
__global__ void __launch_bounds__(256, 2) k_from_x(void* __restrict__ y, uint8_t* __restrict__ ok, const void* __restrict__ x, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Lazy md;
  fe r;
  uint8_t o;
  from_x_compute<QUIRK>(S::load(x, n, i, 1, 0), r, o, md);
  if (__builtin_expect(md.flagged(), 0)) { from_x_exact<QUIRK>(y, ok, x, n, i); return; }
  ok[i] = o;
  S::store(y, n, i, 1, 0, r);
}

__global__ void __launch_bounds__(256) k_from_affine(void* __restrict__ J, const void* __restrict__ xy, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  jac r;
  r.x = fp_from_classical(S::load(xy, n, i, 2, 0));
  r.y = fp_from_classical(S::load(xy, n, i, 2, 1));
  r.z = fe_R();
  store_jac(J, n, i, r);
}

// ---- host-side staging: bring every operand to a device SOA buffer and back --------------
struct Staged {
  Scratch sc;
  cudaStream_t s;
  uint32_t flags;
  size_t n;
  struct Out { void* user; void* dev_soa; void* dev_raw; int nc; };
  std::vector<Out> outs;
  Staged(cudaStream_t st, uint32_t f, size_t nn) : sc(st), s(st), flags(f), n(nn) {}

  // returns a device SOA pointer holding the input
  int in(const void* user, int nc, const void** dev) {
    const int L = layout_of(flags);
    const size_t bytes = operand_bytes(n, nc);
    const void* raw = user;
    int rc;
    if (!on_device(flags)) {
      void* d;
      if ((rc = sc.alloc(&d, bytes))) return rc;
      ECB_CUDA(cudaMemcpyAsync(d, user, bytes, cudaMemcpyHostToDevice, s));
      raw = d;
    }
    if (L == L_SOA) { *dev = raw; return ECB200_OK; }
    void* soa;
    if ((rc = sc.alloc(&soa, bytes))) return rc;
    if ((rc = convert_to_soa(L, soa, raw, n, nc, s))) return rc;
    *dev = soa;
    return ECB200_OK;
  }
  int out(void* user, int nc, void** dev) {
    const int L = layout_of(flags);
    const size_t bytes = operand_bytes(n, nc);
    int rc;
    Out o{user, nullptr, nullptr, nc};
    if (on_device(flags) && L == L_SOA) { o.dev_soa = user; }
    else {
      if ((rc = sc.alloc(&o.dev_soa, bytes))) return rc;
      if (L != L_SOA) {
        if (on_device(flags)) o.dev_raw = user;
        else if ((rc = sc.alloc(&o.dev_raw, bytes))) return rc;
      }
    }
    outs.push_back(o);
    *dev = o.dev_soa;
    return ECB200_OK;
  }
  int finish() {
    const int L = layout_of(flags);
    int rc;
    for (auto& o : outs) {
      const size_t bytes = operand_bytes(n, o.nc);
      const void* src = o.dev_soa;
      if (L != L_SOA) {
        if ((rc = convert_from_soa(L, o.dev_raw, o.dev_soa, n, o.nc, s))) return rc;
        src = o.dev_raw;
      }
      if (!on_device(flags)) ECB_CUDA(cudaMemcpyAsync(o.user, src, bytes, cudaMemcpyDeviceToHost, s));
    }
    if (!on_device(flags)) ECB_CUDA(cudaStreamSynchronize(s));
    return ECB200_OK;
  }
};

template <int OP>
static int point_call(void* out1, void* out2, const void* A, const void* B, size_t n, uint32_t flags, void* stream) {
  int rc = check_common(n, flags);
  if (rc) return rc;
  if (n == 0) return ECB200_OK;
  const bool two_out = OP != PO_ADDZ21, two_in = (OP == PO_ZADDU || OP == PO_ZDAU || OP == PO_ADDZ21);
  if (!out1 || !A || (two_out && !out2) || (two_in && !B)) {
    set_error("null pointer argument");
    return ECB200_ERR_ARG;
  }
  Staged st((cudaStream_t)stream, flags, n);
  const void *dA = nullptr, *dB = nullptr;
  void *d1 = nullptr, *d2 = nullptr;
  if ((rc = st.in(A, 3, &dA))) return rc;
  if (two_in && (rc = st.in(B, 3, &dB))) return rc;
  if ((rc = st.out(out1, 3, &d1))) return rc;
  if (two_out && (rc = st.out(out2, 3, &d2))) return rc;
  constexpr int kThreads = PointThreads<OP>::value;
  const unsigned blocks = (unsigned)((n + kThreads - 1) / kThreads);
  if (quirk_on(flags)) k_point<OP, true><<<blocks, kThreads, 0, st.s>>>(d1, d2, dA, dB, n);
  else k_point<OP, false><<<blocks, kThreads, 0, st.s>>>(d1, d2, dA, dB, n);
  ECB_LAUNCH_CHECK();
  return st.finish();
}

constexpr int kLadderThreads = 512;

// Layouts with a native ladder instance (bit-exact mode only: the NO_QUIRK variants exist for SOA).
static bool ladder_has_layout(int L, bool q) { return L == L_SOA || q; }

// Fixed-base table (BASELINE config 4): 2^kBaseTabW ladder states of G, 160 bytes each (10 MiB:
// resident in the 126 MB L2 while a batch runs), one per device and quirk mode, kept in the device's
// context.  ecb200_init builds them (so that a later ECB200_MEM_DEVICE call never synchronises); a device
// used without ecb200_init builds its table on the first fixed-base call, with one stream synchronisation.
// ECB200_BASE_TABLE=0 in the environment selects the plain ladder with P = G.
constexpr int kBaseTabW = 16;
static bool base_table_enabled() {
  static const bool enabled = [] { const char* e = getenv("ECB200_BASE_TABLE"); return !(e && e[0] == '0'); }();
  return enabled;
}
// caller holds c->mu
static int build_base_table_locked(DeviceCtx* c, bool q, cudaStream_t s) {
  if (c->base_tab[q]) return ECB200_OK;
  uint4* t = nullptr;
  ECB_CUDA(cudaMalloc(&t, ((size_t)1 << kBaseTabW) * 160));
  const unsigned blocks = (1u << kBaseTabW) / 128;
  if (q) k_build_base_table<true><<<blocks, 128, 0, s>>>(t, kBaseTabW);
  else k_build_base_table<false><<<blocks, 128, 0, s>>>(t, kBaseTabW);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);   // other streams may use the table from now on
  if (e != cudaSuccess) {
    cudaFree(t);
    set_error("building the fixed-base table failed: %s", cudaGetErrorString(e));
    return ECB200_ERR_CUDA;
  }
  c->base_tab[q] = t;
  return ECB200_OK;
}
static int base_table(bool q, cudaStream_t s, const uint4** out) {
  *out = nullptr;
  if (!base_table_enabled()) return ECB200_OK;
  DeviceCtx* c = current_ctx();
  if (!c) return ECB200_ERR_CUDA;
  std::lock_guard<std::mutex> lock(c->mu);   // per device: other devices' calls are not held up
  int rc = build_base_table_locked(c, q, s);
  if (rc) return rc;
  *out = c->base_tab[q];
  return ECB200_OK;
}
int prebuild_base_tables(DeviceCtx* c) {
  if (!base_table_enabled()) return ECB200_OK;
  std::lock_guard<std::mutex> lock(c->mu);
  int rc = build_base_table_locked(c, true, nullptr);
  return rc ? rc : build_base_table_locked(c, false, nullptr);
}
// ecb200_shutdown: give the device's tables, pipeline streams and bounce buffers back (re-created on demand)
int release_device_resources(DeviceCtx* c) {
  std::lock_guard<std::mutex> lock(c->mu);
  for (auto& t : c->base_tab) {
    if (t) ECB_CUDA(cudaFree(t));
    t = nullptr;
  }
  for (int i = 0; i < 3; i++) {
    if (c->pipe[i]) ECB_CUDA(cudaStreamDestroy(c->pipe[i]));
    c->pipe[i] = nullptr;
    if (c->slot_dev[i]) ECB_CUDA(cudaFree(c->slot_dev[i]));
    c->slot_dev[i] = nullptr;
    if (c->bounce_in[i]) ECB_CUDA(cudaFreeHost(c->bounce_in[i]));
    if (c->bounce_out[i]) ECB_CUDA(cudaFreeHost(c->bounce_out[i]));
    c->bounce_in[i] = c->bounce_out[i] = nullptr;
  }
  c->bounce_in_bytes = c->bounce_out_bytes = 0;
  return ECB200_OK;
}

// One launch for the whole batch, whatever its size: SMs do not run in step (a few are ~3 % slower
// than the rest), so with dynamic block scheduling the fast ones simply take more blocks; splitting
// a batch into "full waves + a tail launch" was measured and loses that slack at the kernel boundary
// (2^20 lanes: 43.6 ms in one launch, 45.2 ms as 13 full waves + tail).
template <bool Q, int MODE, int L, int TABW>
static int launch_ladder_waves(void* dout, const void* dk, const void* dP, int k_bcast, size_t n, cudaStream_t s, const uint4* tab) {
  const unsigned blocks = (unsigned)((n + kLadderThreads - 1) / kLadderThreads);
  k_scalar_mult_sync<Q, MODE, kLadderThreads, L, TABW><<<blocks, kLadderThreads, 0, s>>>(dout, dk, dP, n, k_bcast, tab);
  ECB_LAUNCH_CHECK();
  return ECB200_OK;
}

template <bool Q, int L>
static int launch_ladder_ql(void* dout, const void* dk, const void* dP, int mode, int k_bcast, size_t n, cudaStream_t s, bool use_table) {
  if (mode == 0) return launch_ladder_waves<Q, 0, L, 0>(dout, dk, dP, k_bcast, n, s, nullptr);
  const uint4* tab = nullptr;
  int rc = use_table ? base_table(Q, s, &tab) : ECB200_OK;
  if (rc) return rc;
  if (tab) return launch_ladder_waves<Q, 1, L, kBaseTabW>(dout, dk, dP, k_bcast, n, s, tab);
  return launch_ladder_waves<Q, 1, L, 0>(dout, dk, dP, k_bcast, n, s, nullptr);
}
static int launch_ladder(int L, void* dout, const void* dk, const void* dP, int mode, int k_bcast, size_t n, bool q, cudaStream_t s, bool use_table = true) {
  if (L == L_SOA) return q ? launch_ladder_ql<true, L_SOA>(dout, dk, dP, mode, k_bcast, n, s, use_table) : launch_ladder_ql<false, L_SOA>(dout, dk, dP, mode, k_bcast, n, s, use_table);
  if (L == L_LANE && q) return launch_ladder_ql<true, L_LANE>(dout, dk, dP, mode, k_bcast, n, s, use_table);
  if (L == L_PACK4 && q) return launch_ladder_ql<true, L_PACK4>(dout, dk, dP, mode, k_bcast, n, s, use_table);
  set_error("internal: no ladder instance for layout %d", L);
  return ECB200_ERR_ARG;
}

// Host-memory batches are cut into chunks of one full wave (148 SMs x 512 lanes) that rotate over the
// device's three pipeline streams: the PCIe copies of one chunk overlap the ladder kernel of another, and
// the partial last wave of a chunk overlaps the next chunk's blocks.
//  * pinned (or registered) caller buffers are copied directly;
//  * pageable caller buffers (std::vector, numpy: what the reference's callers hold) go through pinned
//    bounce buffers owned by the library -- a cudaMemcpyAsync on pageable memory would block the host until
//    the copy has happened, i.e. until the chunk's kernel has finished for the copy out, and serialise the
//    pipeline.  The host copies of slot j happen while the other two slots run on the GPU.
//  * affine = true appends to_affine to each chunk (ecb200_scalar_mult_p256_affine): the Jacobian result never
//    crosses PCIe.
//  * each stream owns the device buffers of its chunk (allocated once per device, 21 MiB per stream): chunk ci + 3
//    reuses the buffers of chunk ci in the same stream, after that chunk's copy out.  Allocating them per chunk from
//    the stream-ordered pool looked equivalent but is not: a block freed in one stream and handed to another makes
//    the second stream wait for the first (the pool's internal dependency), i.e. the copy in of chunk ci + 1 waits
//    for the copy out of chunk ci.
constexpr size_t kChunkLanes = 148 * (size_t)kLadderThreads;
constexpr size_t kSlotK = 0, kSlotP = kChunkLanes * 32, kSlotOut = kChunkLanes * 128, kSlotXY = kChunkLanes * 224, kSlotBytes = kChunkLanes * 288;
static int pipe_resources(DeviceCtx* c, bool bounce) {
  std::lock_guard<std::mutex> lock(c->mu);
  for (int i = 0; i < 3; i++) {
    if (!c->pipe[i]) ECB_CUDA(cudaStreamCreateWithFlags(&c->pipe[i], cudaStreamNonBlocking));
    if (!c->slot_dev[i]) ECB_CUDA(cudaMalloc(&c->slot_dev[i], kSlotBytes));
  }
  if (bounce && !c->bounce_in[0]) {
    c->bounce_in_bytes = kChunkLanes * 128;
    c->bounce_out_bytes = kChunkLanes * 96;
    for (int i = 0; i < 3; i++) {
      ECB_CUDA(cudaHostAlloc(&c->bounce_in[i], c->bounce_in_bytes, cudaHostAllocDefault));
      ECB_CUDA(cudaHostAlloc(&c->bounce_out[i], c->bounce_out_bytes, cudaHostAllocDefault));
    }
  }
  return ECB200_OK;
}
static bool is_pageable(const void* p) {
  if (!p) return false;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return true;
  }
  return a.type == cudaMemoryTypeUnregistered;
}

static int scalar_mult_call(void* out, const void* k, const void* P, int mode, int k_bcast, size_t n, uint32_t flags, void* stream);

static int host_pipeline_body(DeviceCtx* c, void* out, const void* k, const void* P, int mode, size_t n, uint32_t flags, bool affine, bool bounce) {
  const int L = layout_of(flags);
  const bool q = quirk_on(flags);
  const bool native = ladder_has_layout(L, q);
  const bool use_table = (flags & ECB200_NO_BASE_TABLE) == 0;
  const uint32_t dflags = (flags & ~ECB200_MEM_MASK) | ECB200_MEM_DEVICE;
  const size_t out_lane = affine ? 64 : 96;
  cudaStream_t* ss = c->pipe;
  size_t done_lo[3] = {0, 0, 0}, done_m[3] = {0, 0, 0};   // chunk whose results wait in bounce_out[j]
  int rc;
  size_t ci = 0;
  for (size_t lo = 0; lo < n; lo += kChunkLanes, ci++) {
    const size_t m = (n - lo < kChunkLanes) ? n - lo : kChunkLanes;
    const int j = (int)(ci % 3);
    cudaStream_t s = ss[j];
    const char* src_k = (const char*)k + lo * 32;
    const char* src_P = P ? (const char*)P + lo * 96 : nullptr;
    char* dst = (char*)out + lo * out_lane;
    if (bounce) {
      if (ci >= 3) {   // slot j is free again once its previous chunk has left the GPU
        ECB_CUDA(cudaStreamSynchronize(s));
        memcpy((char*)out + done_lo[j] * out_lane, c->bounce_out[j], done_m[j] * out_lane);
      }
      memcpy(c->bounce_in[j], src_k, m * 32);
      src_k = (const char*)c->bounce_in[j];
      if (mode == 0) {
        memcpy((char*)c->bounce_in[j] + kChunkLanes * 32, src_P, m * 96);
        src_P = (const char*)c->bounce_in[j] + kChunkLanes * 32;
      }
      dst = (char*)c->bounce_out[j];
      done_lo[j] = lo;
      done_m[j] = m;
    }
    Scratch sc(s);   // only the layouts without a native ladder instance need more than the slot's own buffers
    sc.ctx = c;
    char* slot = (char*)c->slot_dev[j];
    void *rk = slot + kSlotK, *rP = nullptr, *ro = slot + kSlotOut, *rxy = nullptr;
    // LANE and PACK4 are both contiguous per group of 4 lanes: a chunk is a byte range
    ECB_CUDA(cudaMemcpyAsync(rk, src_k, operand_bytes(m, 1), cudaMemcpyHostToDevice, s));
    if (mode == 0) {
      rP = slot + kSlotP;
      ECB_CUDA(cudaMemcpyAsync(rP, src_P, operand_bytes(m, 3), cudaMemcpyHostToDevice, s));
    }
    if (native) {
      if ((rc = launch_ladder(L, ro, rk, rP, mode, 0, m, q, s, use_table))) return rc;
    } else {
      void *sk, *sP = nullptr, *so;
      if ((rc = sc.alloc(&sk, operand_bytes(m, 1))) || (rc = sc.alloc(&so, operand_bytes(m, 3)))) return rc;
      if ((rc = convert_to_soa(L, sk, rk, m, 1, s))) return rc;
      if (mode == 0) {
        if ((rc = sc.alloc(&sP, operand_bytes(m, 3)))) return rc;
        if ((rc = convert_to_soa(L, sP, rP, m, 3, s))) return rc;
      }
      if ((rc = launch_ladder(L_SOA, so, sk, sP, mode, 0, m, q, s, use_table))) return rc;
      if ((rc = convert_from_soa(L, ro, so, m, 3, s))) return rc;
    }
    const void* res = ro;
    if (affine) {
      rxy = slot + kSlotXY;
      if ((rc = ecb200_to_affine(rxy, ro, m, dflags, s))) return rc;
      res = rxy;
    }
    ECB_CUDA(cudaMemcpyAsync(dst, res, m * out_lane, cudaMemcpyDeviceToHost, s));
  }
  // the last (up to) three chunks, oldest first: the host copy out of one overlaps the GPU work of the next
  for (size_t t = ci >= 3 ? ci - 3 : 0; t < ci; t++) {
    const int j = (int)(t % 3);
    ECB_CUDA(cudaStreamSynchronize(ss[j]));
    if (bounce && done_m[j]) memcpy((char*)out + done_lo[j] * out_lane, c->bounce_out[j], done_m[j] * out_lane);
  }
  for (int j = 0; j < 3; j++) ECB_CUDA(cudaStreamSynchronize(ss[j]));   // fewer than three chunks: idle streams, no-op
  return ECB200_OK;
}

// one device: the calling thread's current one
static int host_pipeline(void* out, const void* k, const void* P, int mode, size_t n, uint32_t flags, bool affine) {
  DeviceCtx* c = current_ctx();
  if (!c) return ECB200_ERR_CUDA;
  const bool bounce = is_pageable(k) || is_pageable(P) || is_pageable(out);
  int rc = pipe_resources(c, bounce);
  if (rc) return rc;
  std::lock_guard<std::mutex> lock(c->pipe_mu);
  rc = host_pipeline_body(c, out, k, P, mode, n, flags, affine, bounce);
  if (rc) {
    // a chunk failed: earlier chunks are still in flight and use the caller's buffers -- wait for them
    // before handing the error back (the message of the first failure is kept)
    char msg[512];
    snprintf(msg, sizeof msg, "%s", last_error());
    for (int j = 0; j < 3; j++) cudaStreamSynchronize(c->pipe[j]);
    cudaGetLastError();
    set_error("%s", msg);
  }
  return rc;
}

// ecb200_init_devices: the batch is cut into one contiguous index range per device (lanes are independent:
// include/ecsimd/bignum.h:101-102), each range on its own host thread through that device's pipeline.
std::vector<int> multi_devices();
static int host_pipeline_multi(const std::vector<int>& devs, void* out, const void* k, const void* P, int mode, size_t n, uint32_t flags, bool affine) {
  const size_t D = devs.size();
  const size_t out_lane = affine ? 64 : 96;
  // ranges in units of 512 lanes (whole thread blocks; also a multiple of the 4-lane packs)
  const size_t units = (n + kLadderThreads - 1) / kLadderThreads;
  std::vector<size_t> lo(D + 1);
  for (size_t d = 0; d <= D; d++) {
    const size_t u = units * d / D;
    lo[d] = (u * kLadderThreads < n) ? u * kLadderThreads : n;
  }
  lo[D] = n;
  std::vector<int> rcs(D, ECB200_OK);
  std::vector<std::string> msgs(D);
  int prev = 0;
  ECB_CUDA(cudaGetDevice(&prev));
  auto work = [&](size_t d) {
    const size_t a = lo[d], m = lo[d + 1] - lo[d];
    if (m == 0) return;
    cudaError_t e = cudaSetDevice(devs[d]);
    if (e != cudaSuccess) {
      rcs[d] = ECB200_ERR_CUDA;
      msgs[d] = std::string("cudaSetDevice failed: ") + cudaGetErrorString(e);
      return;
    }
    rcs[d] = host_pipeline((char*)out + a * out_lane, (const char*)k + a * 32, P ? (const char*)P + a * 96 : nullptr, mode, m, flags, affine);
    if (rcs[d]) msgs[d] = last_error();
  };
  std::vector<std::thread> th;
  for (size_t d = 1; d < D; d++) th.emplace_back(work, d);
  work(0);
  for (auto& t : th) t.join();
  ECB_CUDA(cudaSetDevice(prev));
  for (size_t d = 0; d < D; d++)
    if (rcs[d]) {
      set_error("device %d: %s", devs[d], msgs[d].c_str());
      return rcs[d];
    }
  return ECB200_OK;
}

// host-memory batch of the scalar multiplication (optionally followed by to_affine)
static int host_batch(void* out, const void* k, const void* P, int mode, size_t n, uint32_t flags, bool affine, cudaStream_t user) {
  ECB_CUDA(cudaStreamSynchronize(user));  // host buffers are the caller's: order after its pending work
  const std::vector<int> devs = multi_devices();
  if (devs.size() > 1 && n >= 2 * kChunkLanes) return host_pipeline_multi(devs, out, k, P, mode, n, flags, affine);
  return host_pipeline(out, k, P, mode, n, flags, affine);
}

static int scalar_mult_call(void* out, const void* k, const void* P, int mode, int k_bcast, size_t n, uint32_t flags, void* stream) {
  int rc = check_common(n, flags);
  if (rc) return rc;
  if (n == 0) return ECB200_OK;
  if (!out || !k || (mode == 0 && !P)) {
    set_error("null pointer argument");
    return ECB200_ERR_ARG;
  }
  if (!on_device(flags) && layout_of(flags) != L_SOA && !k_bcast && n > kChunkLanes)
    return host_batch(out, k, P, mode, n, flags, false, (cudaStream_t)stream);
  const int L = layout_of(flags);
  const bool q = quirk_on(flags);
  // native layout: the kernel reads/writes the caller's layout (only the memory space is staged)
  const bool native = ladder_has_layout(L, q);
  Staged st((cudaStream_t)stream, native ? ((flags & ~ECB200_LAYOUT_MASK) | ECB200_LAYOUT_SOA) : flags, n);  // SOA = "no conversion" for Staged
  const void *dk = nullptr, *dP = nullptr;
  void* dout = nullptr;
  if (k_bcast) {
    void* d;
    if ((rc = st.sc.alloc(&d, 32))) return rc;
    ECB_CUDA(cudaMemcpyAsync(d, k, 32, cudaMemcpyHostToDevice, st.s));
    dk = d;
  } else if ((rc = st.in(k, 1, &dk))) return rc;
  if (mode == 0 && (rc = st.in(P, 3, &dP))) return rc;
  if ((rc = st.out(out, 3, &dout))) return rc;
  if ((rc = launch_ladder(native ? L : L_SOA, dout, dk, dP, mode, k_bcast, n, q, st.s, (flags & ECB200_NO_BASE_TABLE) == 0))) return rc;
  return st.finish();
}

}  // namespace ecb200

using namespace ecb200;

extern "C" {

int ecb200_dblu(void* outP, void* out2, const void* P, size_t n, uint32_t flags, void* stream) {
  return point_call<PO_DBLU>(outP, out2, P, nullptr, n, flags, stream);
}
int ecb200_zaddu(void* outP, void* outR, const void* P, const void* O, size_t n, uint32_t flags, void* stream) {
  return point_call<PO_ZADDU>(outP, outR, P, O, n, flags, stream);
}
int ecb200_zdau(void* outQ, void* outR, const void* P, const void* Q, size_t n, uint32_t flags, void* stream) {
  return point_call<PO_ZDAU>(outQ, outR, P, Q, n, flags, stream);
}
int ecb200_add_z2_1(void* outR, const void* A, const void* B, size_t n, uint32_t flags, void* stream) {
  return point_call<PO_ADDZ21>(outR, nullptr, A, B, n, flags, stream);
}
int ecb200_trplu(void* outP, void* out3, const void* P, size_t n, uint32_t flags, void* stream) {
  return point_call<PO_TRPLU>(outP, out3, P, nullptr, n, flags, stream);
}
int ecb200_scalar_mult_p256(void* out, const void* k, const void* P, size_t n, uint32_t flags, void* stream) {
  return scalar_mult_call(out, k, P, 0, 0, n, flags, stream);
}
int ecb200_scalar_mult_p256_base(void* out, const void* k, size_t n, uint32_t flags, void* stream) {
  return scalar_mult_call(out, k, nullptr, 1, 0, n, flags, stream);
}
int ecb200_scalar_mult_p256_1s(void* out, const uint32_t* k1, const void* P, size_t n, uint32_t flags, void* stream) {
  return scalar_mult_call(out, k1, P, 0, 1, n, flags, stream);
}

int ecb200_from_affine(void* outJ, const void* xy, size_t n, uint32_t flags, void* stream) {
  int rc = check_common(n, flags);
  if (rc) return rc;
  if (n == 0) return ECB200_OK;
  if (!outJ || !xy) { set_error("null pointer argument"); return ECB200_ERR_ARG; }
  Staged st((cudaStream_t)stream, flags, n);
  const void* din = nullptr;
  void* dout = nullptr;
  if ((rc = st.in(xy, 2, &din))) return rc;
  if ((rc = st.out(outJ, 3, &dout))) return rc;
  k_from_affine<<<(unsigned)((n + 255) / 256), 256, 0, st.s>>>(dout, din, n);
  ECB_LAUNCH_CHECK();
  return st.finish();
}

int ecb200_from_x(void* y, uint8_t* ok, const void* x, size_t n, uint32_t flags, void* stream) {
  int rc = check_common(n, flags);
  if (rc) return rc;
  if (n == 0) return ECB200_OK;
  if (!y || !ok || !x) { set_error("null pointer argument"); return ECB200_ERR_ARG; }
  Staged st((cudaStream_t)stream, flags, n);
  const void* din = nullptr;
  void* dout = nullptr;
  void* dok = ok;
  if ((rc = st.in(x, 1, &din))) return rc;
  if ((rc = st.out(y, 1, &dout))) return rc;
  if (!on_device(flags) && (rc = st.sc.alloc(&dok, n))) return rc;
  const unsigned blocks = (unsigned)((n + 255) / 256);
  if (quirk_on(flags)) k_from_x<true><<<blocks, 256, 0, st.s>>>(dout, (uint8_t*)dok, din, n);
  else k_from_x<false><<<blocks, 256, 0, st.s>>>(dout, (uint8_t*)dok, din, n);
  ECB_LAUNCH_CHECK();
  if (!on_device(flags)) ECB_CUDA(cudaMemcpyAsync(ok, dok, n, cudaMemcpyDeviceToHost, st.s));
  return st.finish();
}

int ecb200_to_affine(void* xy, const void* J, size_t n, uint32_t flags, void* stream) {
  int rc = check_common(n, flags);
  if (rc) return rc;
  if (n == 0) return ECB200_OK;
  if (!xy || !J) { set_error("null pointer argument"); return ECB200_ERR_ARG; }
  Staged st((cudaStream_t)stream, flags, n);
  const void* din = nullptr;
  void* dout = nullptr;
  if ((rc = st.in(J, 3, &din))) return rc;
  if ((rc = st.out(xy, 2, &dout))) return rc;
  const unsigned blocks = (unsigned)((n + 255) / 256);
  if (quirk_on(flags)) k_to_affine<true><<<blocks, 256, 0, st.s>>>(dout, din, n);
  else k_to_affine<false><<<blocks, 256, 0, st.s>>>(dout, din, n);
  ECB_LAUNCH_CHECK();
  return st.finish();
}

// scalar_mult(k, P).to_affine() in one call (what benchs/curve_group.cpp:28-46 of the reference times):
// the Jacobian result stays on the device, so host callers move 128 B in and 64 B out per lane
// instead of 128 + 96 and then 96 + 64.  Host batches are cut into the same one-wave chunks over
// three streams as ecb200_scalar_mult_p256, each chunk running ladder -> to_affine -> copy out.
static int scalar_mult_affine_call(void* out_xy, const void* k, const void* P, int mode, size_t n, uint32_t flags, void* stream) {
  int rc = check_common(n, flags);
  if (rc) return rc;
  if (n == 0) return ECB200_OK;
  if (!out_xy || !k || (mode == 0 && !P)) {
    set_error("null pointer argument");
    return ECB200_ERR_ARG;
  }
  cudaStream_t user = (cudaStream_t)stream;
  if (on_device(flags)) {
    Scratch sc(user);
    void* J;
    if ((rc = sc.alloc(&J, operand_bytes(n, 3)))) return rc;
    if ((rc = scalar_mult_call(J, k, P, mode, 0, n, flags, stream))) return rc;
    return ecb200_to_affine(out_xy, J, n, flags, stream);
  }
  // LANE and PACK4 are contiguous per group of 4 lanes (a chunk is a byte range): the chunked pipeline
  if (layout_of(flags) != L_SOA) return host_batch(out_xy, k, P, mode, n, flags, true, user);
  // planar host buffers: one pass through device temporaries
  const uint32_t dflags = (flags & ~ECB200_MEM_MASK) | ECB200_MEM_DEVICE;
  ECB_CUDA(cudaStreamSynchronize(user));
  Scratch sc(user);
  void *rk, *rP = nullptr, *rJ, *rxy;
  if ((rc = sc.alloc(&rk, operand_bytes(n, 1))) || (rc = sc.alloc(&rJ, operand_bytes(n, 3))) || (rc = sc.alloc(&rxy, operand_bytes(n, 2)))) return rc;
  ECB_CUDA(cudaMemcpyAsync(rk, k, operand_bytes(n, 1), cudaMemcpyHostToDevice, user));
  if (mode == 0) {
    if ((rc = sc.alloc(&rP, operand_bytes(n, 3)))) return rc;
    ECB_CUDA(cudaMemcpyAsync(rP, P, operand_bytes(n, 3), cudaMemcpyHostToDevice, user));
  }
  if ((rc = scalar_mult_call(rJ, rk, rP, mode, 0, n, dflags, user))) return rc;
  if ((rc = ecb200_to_affine(rxy, rJ, n, dflags, user))) return rc;
  ECB_CUDA(cudaMemcpyAsync(out_xy, rxy, operand_bytes(n, 2), cudaMemcpyDeviceToHost, user));
  ECB_CUDA(cudaStreamSynchronize(user));
  return ECB200_OK;
}

int ecb200_scalar_mult_p256_affine(void* out_xy, const void* k, const void* P, size_t n, uint32_t flags, void* stream) {
  return scalar_mult_affine_call(out_xy, k, P, P ? 0 : 1, n, flags, stream);
}

}  // extern "C"
