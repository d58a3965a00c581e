// layout.cuh -- the three memory layouts of the C ABI (include/ecb200.h) as
// compile-time accessors.  `i` is the lane index, `nc` the number of coordinates
// per lane in the buffer (1 value, 2 affine, 3 Jacobian), `c` the coordinate.
#pragma once
#include <cstddef>
#include <cstdint>

#include "fp256.cuh"

namespace ecb200 {

enum : int { L_LANE = 0, L_PACK4 = 1, L_SOA = 2 };

template <int L>
struct Layout;

// value i = 8 consecutive u32; a thread moves it as two 128-bit accesses
template <>
struct Layout<L_LANE> {
  static __device__ __forceinline__ fe load(const void* base, size_t n, size_t i, int nc, int c) {
    const uint4* p = reinterpret_cast<const uint4*>(base) + (i * nc + c) * 2;
    const uint4 lo = __ldg(p), hi = __ldg(p + 1);
    fe r;
    r.v[0] = lo.x; r.v[1] = lo.y; r.v[2] = lo.z; r.v[3] = lo.w;
    r.v[4] = hi.x; r.v[5] = hi.y; r.v[6] = hi.z; r.v[7] = hi.w;
    return r;
  }
  static __device__ __forceinline__ void store(void* base, size_t n, size_t i, int nc, int c, const fe& a) {
    uint4* p = reinterpret_cast<uint4*>(base) + (i * nc + c) * 2;
    p[0] = make_uint4(a.v[0], a.v[1], a.v[2], a.v[3]);
    p[1] = make_uint4(a.v[4], a.v[5], a.v[6], a.v[7]);
  }
  // 32-bit word w (0..7) of value i
  static __device__ __forceinline__ uint32_t load_word(const void* base, size_t n, size_t i, int nc, int c, int w) {
    return __ldg(reinterpret_cast<const uint32_t*>(base) + (i * nc + c) * 8 + w);
  }
};

// the reference's wide<bignum_256> pack: u64 word index inside a pack = limb*4 + lane
// (include/ecsimd/bignum.h:101-102); multi-coordinate objects are pack|pack|pack.
template <>
struct Layout<L_PACK4> {
  static __device__ __forceinline__ fe load(const void* base, size_t n, size_t i, int nc, int c) {
    const uint2* p = reinterpret_cast<const uint2*>(base) + ((i >> 2) * nc + c) * 16 + (i & 3);
    fe r;
#pragma unroll
    for (int l = 0; l < 4; l++) {
      const uint2 w = __ldg(p + 4 * l);
      r.v[2 * l] = w.x;
      r.v[2 * l + 1] = w.y;
    }
    return r;
  }
  static __device__ __forceinline__ void store(void* base, size_t n, size_t i, int nc, int c, const fe& a) {
    uint2* p = reinterpret_cast<uint2*>(base) + ((i >> 2) * nc + c) * 16 + (i & 3);
#pragma unroll
    for (int l = 0; l < 4; l++) p[4 * l] = make_uint2(a.v[2 * l], a.v[2 * l + 1]);
  }
  static __device__ __forceinline__ uint32_t load_word(const void* base, size_t n, size_t i, int nc, int c, int w) {
    const uint32_t* p = reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint2*>(base) + ((i >> 2) * nc + c) * 16 + (i & 3) + 4 * (w >> 1));
    return __ldg(p + (w & 1));
  }
};

// planar: plane (2c) holds words 0..3 of coordinate c for all lanes, plane (2c+1) words 4..7
template <>
struct Layout<L_SOA> {
  static __device__ __forceinline__ fe load(const void* base, size_t n, size_t i, int nc, int c) {
    const uint4* p = reinterpret_cast<const uint4*>(base) + (size_t)(2 * c) * n + i;
    const uint4 lo = __ldg(p), hi = __ldg(p + n);
    fe r;
    r.v[0] = lo.x; r.v[1] = lo.y; r.v[2] = lo.z; r.v[3] = lo.w;
    r.v[4] = hi.x; r.v[5] = hi.y; r.v[6] = hi.z; r.v[7] = hi.w;
    return r;
  }
  static __device__ __forceinline__ void store(void* base, size_t n, size_t i, int nc, int c, const fe& a) {
    uint4* p = reinterpret_cast<uint4*>(base) + (size_t)(2 * c) * n + i;
    p[0] = make_uint4(a.v[0], a.v[1], a.v[2], a.v[3]);
    p[n] = make_uint4(a.v[4], a.v[5], a.v[6], a.v[7]);
  }
  static __device__ __forceinline__ uint32_t load_word(const void* base, size_t n, size_t i, int nc, int c, int w) {
    const uint32_t* p = reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint4*>(base) + (size_t)(2 * c + (w >> 2)) * n + i);
    return __ldg(p + (w & 3));
  }
};

}  // namespace ecb200
