// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Thin array driver around the UNMODIFIED aguinet/ecsimd headers under
// /root/reference (include/ecsimd + vendored third-party/eve, ctbignum).
// Built by oracle/Makefile into oracle/_ref/libecsimd_ref.so (git-ignored).
// It is the ground truth the C restatement (oracle/p256_oracle.c) is pinned
// against, the generator of tests/golden/*, and the "reference" CPU baseline
// timed by bench.py.  No reference source is copied: this file only *calls*
// the reference's public templates:
//   mgry_add/sub/shift_left/mul/sqr      include/ecsimd/mgry_ops.h:10-42
//   mul / square / mgry_reduce           include/ecsimd/mul.h:150-221, mgry_mul.h:84-121
//   DBLU/ZADDU/ZDAU/ADD_Z2_1/TRPLU       include/ecsimd/curve_group.h:64-186
//   scalar_mult / scalar_mult_1s         include/ecsimd/curve_group.h:189-251
//   from_affine / to_affine / opposite   include/ecsimd/jacobian_curve_point.h:25-54
//   from_x                               include/ecsimd/curve_point_ops.h:12-22
//
// Flat array convention (same as the C oracle and the CUDA C-ABI "lane"
// layout): a 256-bit value is 4 x u64, least-significant limb first; a
// Jacobian point is X|Y|Z = 12 x u64; an affine point is x|y = 8 x u64.
// Lanes are independent; the harness gathers 4 lanes into one ecsimd pack,
// pads a ragged tail with zero lanes and drops their outputs.
#include <ecsimd/curve_group.h>
#include <ecsimd/curve_nist_p256.h>
#include <ecsimd/curve_point.h>
#include <ecsimd/curve_point_ops.h>
#include <ecsimd/jacobian_curve_point.h>
#include <ecsimd/mgry_ops.h>
#include <ecsimd/gfp.h>

#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

using namespace ecsimd;
using Curve = curve_nist_p256;
using CG = curve_group<Curve>;
using BN = bignum_256;
using WBN = curve_wide_bn_t<Curve>;
using WBN512 = wide_bignum<bignum_512>;
using WMBN = curve_wide_mgry_bn_t<Curve>;
using GFP = GFp<WBN, Curve::P>;
using WCP = wide_curve_point<Curve>;
using WJCP = wide_jacobian_curve_point<Curve>;

namespace {

// pack word order is limb*4 + lane (EVE stores wide<struct> as per-field wides)
inline WBN load4(const uint64_t* base, size_t stride, size_t i, size_t n) {
  alignas(32) uint64_t buf[16];
  for (size_t lane = 0; lane < 4; ++lane)
    for (size_t l = 0; l < 4; ++l)
      buf[l * 4 + lane] = (i + lane < n) ? base[(i + lane) * stride + l] : 0;
  WBN w;
  static_assert(sizeof(WBN) == sizeof(buf));
  std::memcpy(&w, buf, sizeof(buf));
  return w;
}
inline void store4(uint64_t* base, size_t stride, size_t i, size_t n, WBN const& w) {
  alignas(32) uint64_t buf[16];
  std::memcpy(buf, &w, sizeof(buf));
  for (size_t lane = 0; lane < 4 && i + lane < n; ++lane)
    for (size_t l = 0; l < 4; ++l)
      base[(i + lane) * stride + l] = buf[l * 4 + lane];
}
inline WBN512 load4_512(const uint64_t* base, size_t i, size_t n) {
  alignas(32) uint64_t buf[32];
  for (size_t lane = 0; lane < 4; ++lane)
    for (size_t l = 0; l < 8; ++l)
      buf[l * 4 + lane] = (i + lane < n) ? base[(i + lane) * 8 + l] : 0;
  WBN512 w;
  static_assert(sizeof(WBN512) == sizeof(buf));
  std::memcpy(&w, buf, sizeof(buf));
  return w;
}
inline void store4_512(uint64_t* base, size_t i, size_t n, WBN512 const& w) {
  alignas(32) uint64_t buf[32];
  std::memcpy(buf, &w, sizeof(buf));
  for (size_t lane = 0; lane < 4 && i + lane < n; ++lane)
    for (size_t l = 0; l < 8; ++l)
      base[(i + lane) * 8 + l] = buf[l * 4 + lane];
}
inline GFP gf(WBN const& w) { return GFP{WMBN{w}}; }
inline WJCP loadJ(const uint64_t* p, size_t i, size_t n) {
  WJCP r;
  r.x() = gf(load4(p, 12, i, n));
  r.y() = gf(load4(p + 4, 12, i, n));
  r.z() = gf(load4(p + 8, 12, i, n));
  return r;
}
inline void storeJ(uint64_t* p, size_t i, size_t n, WJCP const& r) {
  store4(p, 12, i, n, r.x().wbn());
  store4(p + 4, 12, i, n, r.y().wbn());
  store4(p + 8, 12, i, n, r.z().wbn());
}

template <class F>
void par_for4(size_t n, int nthreads, F f) {
  size_t npacks = (n + 3) / 4;
  if (nthreads <= 1 || npacks < 2) {
    for (size_t p = 0; p < npacks; ++p) f(p * 4);
    return;
  }
  std::vector<std::thread> th;
  size_t per = (npacks + nthreads - 1) / nthreads;
  for (int t = 0; t < nthreads; ++t) {
    size_t lo = t * per, hi = std::min(npacks, lo + per);
    if (lo >= hi) break;
    th.emplace_back([=] { for (size_t p = lo; p < hi; ++p) f(p * 4); });
  }
  for (auto& t : th) t.join();
}

}  // namespace

extern "C" {

int ref_abi_version() { return 1; }

// ---- layout probe: lets the tests assert the pack4 word order ------------
void ref_pack_layout_probe(uint64_t* out16, const uint64_t* four_lanes /*4x4*/) {
  WBN w = load4(four_lanes, 4, 0, 4);
  std::memcpy(out16, &w, 128);
}
size_t ref_sizeof_wbn() { return sizeof(WBN); }
size_t ref_sizeof_wjcp() { return sizeof(WJCP); }
size_t ref_sizeof_wcp() { return sizeof(WCP); }

// ---- field ops -------------------------------------------------------------
void ref_mgry_add(uint64_t* o, const uint64_t* a, const uint64_t* b, size_t n, int nt) {
  par_for4(n, nt, [=](size_t i) { store4(o, 4, i, n, mgry_add(WMBN{load4(a, 4, i, n)}, WMBN{load4(b, 4, i, n)}).wbn()); });
}
void ref_mgry_sub(uint64_t* o, const uint64_t* a, const uint64_t* b, size_t n, int nt) {
  par_for4(n, nt, [=](size_t i) { store4(o, 4, i, n, mgry_sub(WMBN{load4(a, 4, i, n)}, WMBN{load4(b, 4, i, n)}).wbn()); });
}
void ref_mgry_shl1(uint64_t* o, const uint64_t* a, size_t n, int nt) {
  par_for4(n, nt, [=](size_t i) { store4(o, 4, i, n, mgry_shift_left<1>(WMBN{load4(a, 4, i, n)}).wbn()); });
}
void ref_mgry_mul(uint64_t* o, const uint64_t* a, const uint64_t* b, size_t n, int nt) {
  par_for4(n, nt, [=](size_t i) { store4(o, 4, i, n, mgry_mul(WMBN{load4(a, 4, i, n)}, WMBN{load4(b, 4, i, n)}).wbn()); });
}
void ref_mgry_sqr(uint64_t* o, const uint64_t* a, size_t n, int nt) {
  par_for4(n, nt, [=](size_t i) { store4(o, 4, i, n, mgry_sqr(WMBN{load4(a, 4, i, n)}).wbn()); });
}
void ref_opposite(uint64_t* o, const uint64_t* a, size_t n, int nt) {
  par_for4(n, nt, [=](size_t i) { store4(o, 4, i, n, gf(load4(a, 4, i, n)).opposite().wbn()); });
}
void ref_from_classical(uint64_t* o, const uint64_t* a, size_t n, int nt) {
  par_for4(n, nt, [=](size_t i) { store4(o, 4, i, n, WMBN::from_classical(load4(a, 4, i, n)).wbn()); });
}
void ref_to_classical(uint64_t* o, const uint64_t* a, size_t n, int nt) {
  par_for4(n, nt, [=](size_t i) { store4(o, 4, i, n, WMBN{load4(a, 4, i, n)}.to_classical()); });
}
void ref_inverse(uint64_t* o, const uint64_t* a, size_t n, int nt) {
  par_for4(n, nt, [=](size_t i) { store4(o, 4, i, n, gf(load4(a, 4, i, n)).inverse().wbn()); });
}
// raw integer layer (include/ecsimd/mul.h, mgry_mul.h)
void ref_mul512(uint64_t* o, const uint64_t* a, const uint64_t* b, size_t n) {
  for (size_t i = 0; i < n; i += 4) store4_512(o, i, n, mul(load4(a, 4, i, n), load4(b, 4, i, n)));
}
void ref_square512(uint64_t* o, const uint64_t* a, size_t n) {
  for (size_t i = 0; i < n; i += 4) store4_512(o, i, n, square(load4(a, 4, i, n)));
}
void ref_mgry_reduce(uint64_t* o, const uint64_t* a512, size_t n) {
  for (size_t i = 0; i < n; i += 4) store4(o, 4, i, n, details::mgry_reduce<Curve::P>(load4_512(a512, i, n)));
}

// ---- point ops (Jacobian, Montgomery form; 12 x u64 per lane) -------------
// DBLU: returns 2P in out2, the rewritten P in outP.  curve_group.h:64-87
void ref_dblu(uint64_t* outP, uint64_t* out2, const uint64_t* P, size_t n, int nt) {
  par_for4(n, nt, [=](size_t i) { WJCP p = loadJ(P, i, n); WJCP r = CG::DBLU(p); storeJ(outP, i, n, p); storeJ(out2, i, n, r); });
}
// ZADDU(P&, O): returns P+O in outR, rewritten P in outP.  curve_group.h:91-116
void ref_zaddu(uint64_t* outP, uint64_t* outR, const uint64_t* P, const uint64_t* O, size_t n, int nt) {
  par_for4(n, nt, [=](size_t i) { WJCP p = loadJ(P, i, n); WJCP o = loadJ(O, i, n); WJCP r = CG::ZADDU(p, o); storeJ(outP, i, n, p); storeJ(outR, i, n, r); });
}
// ZDAU(P, Q&): returns 2P+Q in outR, rewritten Q in outQ.  curve_group.h:120-153
void ref_zdau(uint64_t* outQ, uint64_t* outR, const uint64_t* P, const uint64_t* Q, size_t n, int nt) {
  par_for4(n, nt, [=](size_t i) { WJCP p = loadJ(P, i, n); WJCP q = loadJ(Q, i, n); WJCP r = CG::ZDAU(p, q); storeJ(outQ, i, n, q); storeJ(outR, i, n, r); });
}
// ADD_Z2_1(A, B): curve_group.h:155-179
void ref_add_z2_1(uint64_t* outR, const uint64_t* A, const uint64_t* B, size_t n, int nt) {
  par_for4(n, nt, [=](size_t i) { storeJ(outR, i, n, CG::ADD_Z2_1(loadJ(A, i, n), loadJ(B, i, n))); });
}
// TRPLU(P&): curve_group.h:183-186
void ref_trplu(uint64_t* outP, uint64_t* out3, const uint64_t* P, size_t n, int nt) {
  par_for4(n, nt, [=](size_t i) { WJCP p = loadJ(P, i, n); WJCP r = CG::TRPLU(p); storeJ(outP, i, n, p); storeJ(out3, i, n, r); });
}
// scalar_mult(x, P): curve_group.h:189-218 (same body as lib/scalar_mult_p256.cpp:12-14)
void ref_scalar_mult(uint64_t* out, const uint64_t* k, const uint64_t* P, size_t n, int nt) {
  par_for4(n, nt, [=](size_t i) { storeJ(out, i, n, CG::scalar_mult(load4(k, 4, i, n), loadJ(P, i, n))); });
}
// scalar_mult_1s(x, P): one scalar for all lanes. curve_group.h:221-251
void ref_scalar_mult_1s(uint64_t* out, const uint64_t* k1, const uint64_t* P, size_t n, int nt) {
  BN kk; std::memcpy(&kk, k1, 32);
  par_for4(n, nt, [=](size_t i) { storeJ(out, i, n, CG::scalar_mult_1s(kk, loadJ(P, i, n))); });
}
// from_affine: classical (x,y) -> Jacobian-Montgomery, Z = R.  jacobian_curve_point.h:25-31
void ref_from_affine(uint64_t* outJ, const uint64_t* xy, size_t n, int nt) {
  par_for4(n, nt, [=](size_t i) { storeJ(outJ, i, n, WJCP::from_affine(WCP{load4(xy, 8, i, n), load4(xy + 4, 8, i, n)})); });
}
// to_affine: jacobian_curve_point.h:33-42
void ref_to_affine(uint64_t* xy, const uint64_t* J, size_t n, int nt) {
  par_for4(n, nt, [=](size_t i) { WCP a = loadJ(J, i, n).to_affine(); store4(xy, 8, i, n, a.x()); store4(xy + 4, 8, i, n, a.y()); });
}
// from_x: curve_point_ops.h:12-22. ok4[pack] = 1 iff all 4 lanes of the pack decompress (gfp.h:46-54)
void ref_from_x(uint64_t* y, uint8_t* ok4, const uint64_t* x, size_t n) {
  for (size_t i = 0; i < n; i += 4) {
    auto r = WCP::from_x(load4(x, 4, i, n));
    ok4[i / 4] = r ? 1 : 0;
    if (r) store4(y, 4, i, n, r->y());
  }
}
// ---- the reference's field layer on the secp256k1 prime, as in its own tests (tests/mgry.cpp:25-27,
// tests/ops.cpp:221-252).  op codes as in orc_gen_op.
namespace {
struct PK1 {
  static constexpr auto value = bn_from_bytes_BE<bignum_256>("FFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEFFFFFC2F"_hex);
};
using WMBNK = wide_mgry_bignum<WBN, PK1>;
using GFPK = GFp<WBN, PK1>;
}  // namespace
void ref_k1_op(int op, uint64_t* o, const uint64_t* a, const uint64_t* b, const uint64_t* e, size_t n) {
  const WBN p{PK1::value};
  BN ee{};
  if (e) std::memcpy(&ee, e, 32);
  for (size_t i = 0; i < n; i += 4) {
    const WBN x = load4(a, 4, i, n);
    const WBN y = b ? load4(b, 4, i, n) : x;
    WBN r;
    switch (op) {
      case 0: r = mod_add(x, y, p); break;
      case 1: r = mod_sub(x, y, p); break;
      case 2: r = mod_shift_left_one(x, p); break;
      case 3: r = mgry_mul(WMBNK{x}, WMBNK{y}).wbn(); break;
      case 4: r = mgry_sqr(WMBNK{x}).wbn(); break;
      case 5: r = WMBNK::from_classical(x).wbn(); break;
      case 6: r = WMBNK{x}.to_classical(); break;
      case 7: r = mgry_pow(WMBNK{x}, ee).wbn(); break;
      default: r = GFPK{WMBNK{x}}.opposite().wbn(); break;
    }
    store4(o, 4, i, n, r);
  }
}

// constants, for cross-checking the literals baked into the CUDA / C code
void ref_constants(uint64_t* out /* 32 x u64: P, R, R^2, (p-1)R, Am, Bm, Gx_m, Gy_m */) {
  auto put = [&](int idx, BN const& b) { std::memcpy(out + 4 * idx, &b, 32); };
  using C = mgry_constants<WBN, Curve::P>;
  put(0, Curve::P::value); put(1, C::R_p); put(2, C::Rsq_p); put(3, C::Pm1_by_R_p);
  put(4, CG::Am); put(5, CG::Bm);
  WJCP g = CG::WJG(); uint64_t tmp[12]; storeJ(tmp, 0, 1, g); std::memcpy(out + 24, tmp, 64);
}

}  // extern "C"
