// ecsimd.hpp -- drop-in C++ surface of aguinet/ecsimd's P-256 hot path over the B200 engine.
//
// Same names, argument meaning and (absence of) error behaviour as the reference headers,
// re-implemented as thin POD wrappers around the C ABI of include/ecb200.h (the reference's
// EVE-based templates cannot be compiled by nvcc and are not needed):
//
//   ecsimd::bignum_256, wide_bignum<bignum_256>        include/ecsimd/bignum.h:38-102
//   ecsimd::wide_mgry_bignum, mgry_add/sub/mul/sqr/..   include/ecsimd/mgry.h:28-66, mgry_ops.h:10-101
//   ecsimd::GFp (+ operator+,-,*, gfp_shift_left, sqr,  include/ecsimd/gfp.h:17-115
//               inverse, opposite)
//   ecsimd::wide_curve_point, wide_jacobian_curve_point include/ecsimd/curve_point.h:13-43,
//                                                       jacobian_curve_point.h:12-68
//   ecsimd::curve_group<curve_nist_p256>::{DBLU, ZADDU, ZDAU, ADD_Z2_1, TRPLU, scalar_mult,
//               scalar_mult_1s, WG, WJG}                include/ecsimd/curve_group.h:21-252
//   scalar_mult_p256(WBN const&, WJCP const&)           lib/scalar_mult_p256.cpp:12-14
//
// Object layout is the reference's: a wide_bignum is 128 bytes, 32-byte aligned, u64 word index
// limb*4 + lane; a wide_jacobian_curve_point is x|y|z = 384 bytes.  Arrays of these objects can
// therefore be handed to the batch overloads (and to the C ABI with ECB200_LAYOUT_PACK4) as is.
//
// The 4-lane calls exist for source compatibility; throughput comes from the batch overloads
// (`std::size_t npacks` packs per call), which is how a B200 wants to be fed.
//
// Comparisons are per lane, as in the reference (`a == b` is a 4-lane mask, cmp_res_t, bignum.h:136-137;
// tests reduce it with eve::all): wide_mask below plays eve::logical<eve::wide<u64, fixed<4>>>.  The
// directory include/ecsimd_b200/compat/ maps the reference's own header names (<ecsimd/curve_group.h>,
// <eve/function/all.hpp>, <gtest/gtest.h> ...) onto this file, so that the reference's test translation
// units compile UNMODIFIED against the engine (oracle/Makefile: ref_tests_on_b200).
#pragma once
#include <array>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <optional>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "../ecb200.h"

namespace ecsimd {

namespace detail {
inline void check(int rc) {
  // the reference has no error channel at all; a failing CUDA call is the one thing we cannot
  // express as "deterministic garbage", so it throws
  if (rc != ECB200_OK) throw std::runtime_error(std::string("ecb200: ") + ecb200_last_error());
}
constexpr uint32_t kHostPack = ECB200_LAYOUT_PACK4 | ECB200_MEM_HOST;
}  // namespace detail

// ---- per-lane masks (bignum.h:136-137 cmp_res_t; utility.h:44-51 wide_mask_bit) ----------------------
struct alignas(32) wide_mask {
  uint64_t m[4];  // all-ones / zero per lane, the memory image of an AVX2 logical
  static wide_mask splat(bool v) { const uint64_t x = v ? ~uint64_t(0) : 0; return wide_mask{{x, x, x, x}}; }
  bool get(int lane) const { return m[lane] != 0; }
  void set(int lane, bool v) { m[lane] = v ? ~uint64_t(0) : 0; }
  friend wide_mask operator!(wide_mask a) { for (auto& x : a.m) x = ~x; return a; }
  friend wide_mask operator&&(wide_mask a, wide_mask const& b) { for (int k = 0; k < 4; k++) a.m[k] &= b.m[k]; return a; }
  friend wide_mask operator||(wide_mask a, wide_mask const& b) { for (int k = 0; k < 4; k++) a.m[k] |= b.m[k]; return a; }
};
inline bool all(wide_mask const& a) { return a.get(0) && a.get(1) && a.get(2) && a.get(3); }
inline bool any(wide_mask const& a) { return a.get(0) || a.get(1) || a.get(2) || a.get(3); }
inline bool all(bool v) { return v; }
inline bool any(bool v) { return v; }

// ---- bignum.h ------------------------------------------------------------------------------
struct bignum_256 {
  using limb_type = uint64_t;
  static constexpr std::size_t nlimbs = 4;
  uint64_t limb[4];  // least-significant first
  static bignum_256 from(uint64_t v0) { return bignum_256{{v0, 0, 0, 0}}; }
  friend bool operator==(bignum_256 const& a, bignum_256 const& b) { return std::memcmp(&a, &b, sizeof a) == 0; }
  friend bool operator!=(bignum_256 const& a, bignum_256 const& b) { return !(a == b); }
  // cmp.h:11-29 on one value: most significant limb first
  friend bool operator<(bignum_256 const& a, bignum_256 const& b) {
    for (int l = 3; l >= 0; l--) if (a.limb[l] != b.limb[l]) return a.limb[l] < b.limb[l];
    return false;
  }
  friend bool operator>(bignum_256 const& a, bignum_256 const& b) { return b < a; }
  friend bool operator<=(bignum_256 const& a, bignum_256 const& b) { return !(b < a); }
  friend bool operator>=(bignum_256 const& a, bignum_256 const& b) { return !(a < b); }
};

// serialization.h:12-48 (big-endian 32-byte strings) and literals.h:28-43
inline bignum_256 bn_from_bytes_BE(const uint8_t* b) {
  bignum_256 r{};
  for (int i = 0; i < 32; i++) r.limb[(31 - i) / 8] |= uint64_t(b[i]) << (8 * ((31 - i) % 8));
  return r;
}
inline void bn_to_bytes_BE(uint8_t* out, bignum_256 const& v) {
  for (int i = 0; i < 32; i++) out[i] = uint8_t(v.limb[(31 - i) / 8] >> (8 * ((31 - i) % 8)));
}
inline bignum_256 bn_from_hex(const char* hex64) {
  uint8_t b[32];
  auto nib = [](char c) -> int { return c <= '9' ? c - '0' : (c | 32) - 'a' + 10; };
  for (int i = 0; i < 32; i++) b[i] = uint8_t(nib(hex64[2 * i]) << 4 | nib(hex64[2 * i + 1]));
  return bn_from_bytes_BE(b);
}

// the reference's spellings: bn_from_bytes_BE<BN>(ptr | std::array), bn_to_bytes_BE(v) -> std::array
template <class BN>
inline BN bn_from_bytes_BE(const uint8_t* b) { static_assert(std::is_same<BN, bignum_256>::value, "256-bit values only"); return bn_from_bytes_BE(b); }
template <class BN>
inline BN bn_from_bytes_BE(std::array<uint8_t, sizeof(BN)> const& b) { return bn_from_bytes_BE<BN>(b.data()); }
inline std::array<uint8_t, 32> bn_to_bytes_BE(bignum_256 const& v) { std::array<uint8_t, 32> r; bn_to_bytes_BE(r.data(), v); return r; }

namespace literals {
// "6b17..."_hex -> std::array<uint8_t, len/2>   (literals.h:28-43; the string-literal operator template is the
// same GNU extension the reference relies on)
#if defined(__GNUC__)
#pragma GCC diagnostic push
#pragma GCC diagnostic ignored "-Wpedantic"
#if defined(__clang__)
#pragma GCC diagnostic ignored "-Wgnu-string-literal-operator-template"
#endif
template <class CharT, CharT... Str>
constexpr auto operator"" _hex() {
  static_assert(std::is_same<CharT, char>::value && sizeof...(Str) % 2 == 0, "an even number of hexadecimal digits");
  constexpr char txt[] = {Str..., 0};
  std::array<uint8_t, sizeof...(Str) / 2> out{};
  for (std::size_t i = 0; i < out.size(); i++) {
    uint8_t v = 0;
    for (int h = 0; h < 2; h++) {
      const char c = txt[2 * i + h];
      v = uint8_t(v << 4 | (c >= '0' && c <= '9' ? c - '0' : (c | 32) - 'a' + 10));
    }
    out[i] = v;
  }
  return out;
}
#pragma GCC diagnostic pop
#endif
}  // namespace literals

template <class BN>
struct wide_bignum;
template <>
struct alignas(32) wide_bignum<bignum_256> {
  using value_type = bignum_256;
  static constexpr std::size_t cardinal = 4;
  uint64_t w[16];  // word index = limb*4 + lane
  wide_bignum() = default;
  explicit wide_bignum(bignum_256 const& v) { for (int l = 0; l < 4; l++) for (int k = 0; k < 4; k++) w[l * 4 + k] = v.limb[l]; }
  template <class F, class = decltype(std::declval<F>()(0, 4))>
  explicit wide_bignum(F&& gen) { for (int k = 0; k < 4; k++) set(k, gen(k, 4)); }
  bignum_256 get(int lane) const { return bignum_256{{w[lane], w[4 + lane], w[8 + lane], w[12 + lane]}}; }
  void set(int lane, bignum_256 const& v) { for (int l = 0; l < 4; l++) w[l * 4 + lane] = v.limb[l]; }
  // per-lane comparisons (cmp.h:11-29)
  friend wide_mask operator==(wide_bignum const& a, wide_bignum const& b) { wide_mask r; for (int k = 0; k < 4; k++) r.set(k, a.get(k) == b.get(k)); return r; }
  friend wide_mask operator!=(wide_bignum const& a, wide_bignum const& b) { return !(a == b); }
  friend wide_mask operator<(wide_bignum const& a, wide_bignum const& b) { wide_mask r; for (int k = 0; k < 4; k++) r.set(k, a.get(k) < b.get(k)); return r; }
  friend wide_mask operator>(wide_bignum const& a, wide_bignum const& b) { return b < a; }
  friend wide_mask operator<=(wide_bignum const& a, wide_bignum const& b) { return !(b < a); }
  friend wide_mask operator>=(wide_bignum const& a, wide_bignum const& b) { return !(a < b); }
};
using WBN256 = wide_bignum<bignum_256>;
static_assert(sizeof(WBN256) == 128 && alignof(WBN256) == 32, "must match eve::wide<bignum_256, fixed<4>>");
template <class WBN>
using cmp_res_t = wide_mask;
// bit B of each lane's 64-bit word as a mask (utility.h:44-51): the ladder's swap conditions
inline wide_mask wide_mask_bit(const uint64_t (&lanes)[4], unsigned B) { wide_mask r; for (int k = 0; k < 4; k++) r.set(k, (lanes[k] >> B) & 1u); return r; }

// masked select and swap on packs (ifelse.h:15-22, swap.h:15-23): lane k of the result is a[k] where mask[k], else b[k]
inline WBN256 if_else(wide_mask const& mask, WBN256 const& a, WBN256 const& b) {
  WBN256 r;
  for (int l = 0; l < 4; l++) for (int k = 0; k < 4; k++) r.w[l * 4 + k] = mask.get(k) ? a.w[l * 4 + k] : b.w[l * 4 + k];
  return r;
}
inline void swap_if(wide_mask const& mask, WBN256& a, WBN256& b) {
  for (int l = 0; l < 4; l++) for (int k = 0; k < 4; k++) if (mask.get(k)) { const uint64_t t = a.w[l * 4 + k]; a.w[l * 4 + k] = b.w[l * 4 + k]; b.w[l * 4 + k] = t; }
}

struct curve_nist_p256 {  // curve_nist_p256.h:14-32
  using bn_type = bignum_256;
  static bignum_256 P() { return bn_from_hex("ffffffff00000001000000000000000000000000ffffffffffffffffffffffff"); }
  static bignum_256 A() { return bn_from_hex("ffffffff00000001000000000000000000000000fffffffffffffffffffffffc"); }
  static bignum_256 B() { return bn_from_hex("5ac635d8aa3a93e7b3ebbd55769886bc651d06b0cc53b0f63bce3c3e27d2604b"); }
  static bignum_256 Gx() { return bn_from_hex("6b17d1f2e12c4247f8bce6e563a440f277037d812deb33a0f4a13945d898c296"); }
  static bignum_256 Gy() { return bn_from_hex("4fe342e2fe1a7f9b8ee7eb4a7c0f9e162bce33576b315ececbb6406837bf51f5"); }
};
template <class Curve>
using curve_bn_t = typename Curve::bn_type;               // curve.h:25-32
template <class Curve>
using curve_wide_bn_t = wide_bignum<curve_bn_t<Curve>>;
#if __cplusplus >= 202002L
namespace concepts {  // bignum.h:104-112, curve.h:12-21, mgry.h:68-72, gfp.h:92-95 -- satisfied by the one instantiation kept here
template <class T>
concept bignum = std::is_same_v<T, bignum_256>;
template <class T>
concept wide_bignum = std::is_same_v<T, ::ecsimd::wide_bignum<bignum_256>>;
template <class T>
concept curve = std::is_same_v<T, curve_nist_p256>;
template <class T>
concept wst_curve_am3 = curve<T>;
}  // namespace concepts
#endif

// ---- mgry.h / mgry_ops.h ----------------------------------------------------------------------
template <class WBN = WBN256, class P = curve_nist_p256>
struct wide_mgry_bignum {
  using wide_bignum_type = WBN;
  using bignum_type = typename WBN::value_type;
  using P_type = P;
  WBN n_;
  wide_mgry_bignum() = default;
  wide_mgry_bignum(WBN const& n) : n_(n) {}
  static wide_mgry_bignum R() { return wide_mgry_bignum{WBN{bn_from_hex("00000000fffffffeffffffffffffffffffffffff000000000000000000000001")}}; }
  static wide_mgry_bignum from_classical(WBN const& n) { wide_mgry_bignum r; detail::check(ecb200_from_classical(&r.n_, &n, 4, detail::kHostPack, nullptr)); return r; }
  WBN to_classical() const { WBN r; detail::check(ecb200_to_classical(&r, &n_, 4, detail::kHostPack, nullptr)); return r; }
  WBN const& wbn() const { return n_; }
  WBN& wbn() { return n_; }
};
using WMBN = wide_mgry_bignum<>;
static_assert(sizeof(WMBN) == 128, "layout");

inline WMBN mgry_add(WMBN const& a, WMBN const& b) { WMBN r; detail::check(ecb200_mgry_add(&r, &a, &b, 4, detail::kHostPack, nullptr)); return r; }
inline WMBN mgry_sub(WMBN const& a, WMBN const& b) { WMBN r; detail::check(ecb200_mgry_sub(&r, &a, &b, 4, detail::kHostPack, nullptr)); return r; }
inline WMBN mgry_mul(WMBN const& a, WMBN const& b) { WMBN r; detail::check(ecb200_mgry_mul(&r, &a, &b, 4, detail::kHostPack, nullptr)); return r; }
inline WMBN mgry_sqr(WMBN const& a) { WMBN r; detail::check(ecb200_mgry_sqr(&r, &a, 4, detail::kHostPack, nullptr)); return r; }
template <std::size_t Count>
inline WMBN mgry_shift_left(WMBN const& a) { static_assert(Count > 0 && Count <= 8); WMBN r; detail::check(ecb200_mgry_shift_left(&r, &a, int(Count), 4, detail::kHostPack, nullptr)); return r; }
// a**M * R mod p through the reference's LSB-first square-and-multiply (mgry_ops.h:44-86: same sequence of
// squarings, so the same squaring-defect lanes); the exponent is a plain integer, not secret
inline WMBN mgry_pow(WMBN const& a, bignum_256 const& M) {
  WMBN r;
  const bignum_256 p = curve_nist_p256::P();
  detail::check(ecb200_gen_mgry_pow(&r, &a, reinterpret_cast<const uint32_t*>(&M), reinterpret_cast<const uint32_t*>(&p), 4, detail::kHostPack, nullptr));
  return r;
}
inline wide_mask operator==(WMBN const& a, WMBN const& b) { return a.n_ == b.n_; }
inline wide_mask operator!=(WMBN const& a, WMBN const& b) { return a.n_ != b.n_; }
inline WMBN if_else(wide_mask const& mask, WMBN const& a, WMBN const& b) { return WMBN{if_else(mask, a.n_, b.n_)}; }   // ifelse.h:24-28
inline void swap_if(wide_mask const& mask, WMBN& a, WMBN& b) { swap_if(mask, a.n_, b.n_); }                           // swap.h:25-29
inline WMBN operator+(WMBN const& a, WMBN const& b) { return mgry_add(a, b); }
inline WMBN operator-(WMBN const& a, WMBN const& b) { return mgry_sub(a, b); }
inline WMBN operator*(WMBN const& a, WMBN const& b) { return mgry_mul(a, b); }

// batch forms: npacks consecutive 4-lane packs per call
inline void mgry_add(WMBN* out, WMBN const* a, WMBN const* b, std::size_t npacks) { detail::check(ecb200_mgry_add(out, a, b, 4 * npacks, detail::kHostPack, nullptr)); }
inline void mgry_sub(WMBN* out, WMBN const* a, WMBN const* b, std::size_t npacks) { detail::check(ecb200_mgry_sub(out, a, b, 4 * npacks, detail::kHostPack, nullptr)); }
inline void mgry_mul(WMBN* out, WMBN const* a, WMBN const* b, std::size_t npacks) { detail::check(ecb200_mgry_mul(out, a, b, 4 * npacks, detail::kHostPack, nullptr)); }
inline void mgry_sqr(WMBN* out, WMBN const* a, std::size_t npacks) { detail::check(ecb200_mgry_sqr(out, a, 4 * npacks, detail::kHostPack, nullptr)); }

// ---- gfp.h ----------------------------------------------------------------------------------------
template <class WBN_ = WBN256, class P = curve_nist_p256>
struct GFp {
  using WBN = WBN_;
  using BN = typename WBN::value_type;
  using P_type = P;
  WMBN n_;
  GFp() = default;
  GFp(WMBN const& n) : n_(n) {}
  static GFp one() { return GFp{WMBN::R()}; }
  static GFp from_classical(WBN const& n) { return GFp{WMBN::from_classical(n)}; }
  WBN to_classical() const { return n_.to_classical(); }
  GFp inverse() const { GFp r; detail::check(ecb200_gfp_inverse(&r, this, 4, detail::kHostPack, nullptr)); return r; }
  GFp sqr() const { return GFp{mgry_sqr(n_)}; }
  GFp opposite() const { GFp r; detail::check(ecb200_gfp_opposite(&r, this, 4, detail::kHostPack, nullptr)); return r; }
  // sqrt = pow((p+1)/4), valid only if ALL four lanes are squares (gfp.h:46-54)
  std::optional<GFp> sqrt() const {
    const GFp r{mgry_pow(n_, bn_from_hex("3fffffffc0000000400000000000000000000000400000000000000000000000"))};
    if (any(r.sqr().wbn() != wbn())) return {};
    return {r};
  }
  friend wide_mask operator==(GFp const& a, GFp const& b) { return a.n_ == b.n_; }
  friend wide_mask operator!=(GFp const& a, GFp const& b) { return a.n_ != b.n_; }
  WMBN& wmbn() { return n_; }
  WBN const& wbn() const { return n_.wbn(); }
  WBN& wbn() { return n_.wbn(); }
  WMBN const& wmbn() const { return n_; }
};
using gfp_p256 = GFp<>;
static_assert(sizeof(gfp_p256) == 128, "layout");
inline gfp_p256 operator+(gfp_p256 const& a, gfp_p256 const& b) { return gfp_p256{mgry_add(a.n_, b.n_)}; }
inline gfp_p256 operator-(gfp_p256 const& a, gfp_p256 const& b) { return gfp_p256{mgry_sub(a.n_, b.n_)}; }
inline gfp_p256 operator*(gfp_p256 const& a, gfp_p256 const& b) { return gfp_p256{mgry_mul(a.n_, b.n_)}; }
template <std::size_t Count>
inline gfp_p256 gfp_shift_left(gfp_p256 const& a) { return gfp_p256{mgry_shift_left<Count>(a.n_)}; }
inline gfp_p256 if_else(wide_mask const& mask, gfp_p256 const& a, gfp_p256 const& b) { return gfp_p256{if_else(mask, a.n_, b.n_)}; }   // ifelse.h:30-34
inline void swap_if(wide_mask const& mask, gfp_p256& a, gfp_p256& b) { swap_if(mask, a.n_, b.n_); }                                    // swap.h:31-35

// ---- curve_point.h / jacobian_curve_point.h -----------------------------------------------------------
template <class Curve = curve_nist_p256>
struct wide_curve_point {
  using WBN = WBN256;
  WBN x_, y_;
  wide_curve_point() = default;
  wide_curve_point(WBN const& x, WBN const& y) : x_(x), y_(y) {}
  WBN const& x() const { return x_; }
  WBN const& y() const { return y_; }
  WBN& x() { return x_; }
  WBN& y() { return y_; }
  // decompression: y = sqrt(x^3 - 3x + b), all four lanes or nothing (curve_point_ops.h:12-22, curve_group.h:43-58)
  static std::optional<wide_curve_point> from_x(WBN const& x) {
    wide_curve_point r;
    uint8_t ok[4];
    detail::check(ecb200_from_x(&r.y_, ok, &x, 4, detail::kHostPack, nullptr));
    if (!(ok[0] && ok[1] && ok[2] && ok[3])) return {};
    r.x_ = x;
    return {r};
  }
  friend wide_mask operator==(wide_curve_point const& a, wide_curve_point const& b) { return a.x_ == b.x_ && a.y_ == b.y_; }
};
static_assert(sizeof(wide_curve_point<>) == 256, "layout");

template <class Curve = curve_nist_p256>
struct wide_jacobian_curve_point {
  using curve_type = Curve;
  using bignum_type = curve_bn_t<Curve>;
  using WBN = curve_wide_bn_t<Curve>;
  using wide_curve_point_t = wide_curve_point<Curve>;
  using gfp = gfp_p256;
  gfp x_, y_, z_;
  wide_jacobian_curve_point() = default;
  static wide_jacobian_curve_point from_affine(wide_curve_point<Curve> const& pt) {
    wide_jacobian_curve_point r;
    detail::check(ecb200_from_affine(&r, &pt, 4, detail::kHostPack, nullptr));
    return r;
  }
  wide_curve_point<Curve> to_affine() const {
    wide_curve_point<Curve> r;
    detail::check(ecb200_to_affine(&r, this, 4, detail::kHostPack, nullptr));
    return r;
  }
  wide_jacobian_curve_point opposite() const { wide_jacobian_curve_point r = *this; r.y_ = y_.opposite(); return r; }
  gfp& x() { return x_; }
  gfp& y() { return y_; }
  gfp& z() { return z_; }
  gfp const& x() const { return x_; }
  gfp const& y() const { return y_; }
  gfp const& z() const { return z_; }
  friend wide_mask operator==(wide_jacobian_curve_point const& a, wide_jacobian_curve_point const& b) {
    return a.x_.wbn() == b.x_.wbn() && a.y_.wbn() == b.y_.wbn() && a.z_.wbn() == b.z_.wbn();
  }
};
// masked select / swap of points (ifelse.h:36-49, swap.h:37-56); swap_if_same_z leaves Z alone: the two
// points of the co-Z ladder share it
template <class Curve>
inline wide_jacobian_curve_point<Curve> if_else(wide_mask const& mask, wide_jacobian_curve_point<Curve> const& A, wide_jacobian_curve_point<Curve> const& B) {
  wide_jacobian_curve_point<Curve> r;
  r.x() = if_else(mask, A.x(), B.x());
  r.y() = if_else(mask, A.y(), B.y());
  r.z() = if_else(mask, A.z(), B.z());
  return r;
}
template <class Curve>
inline void swap_if(wide_mask const& mask, wide_jacobian_curve_point<Curve>& A, wide_jacobian_curve_point<Curve>& B) {
  swap_if(mask, A.x(), B.x());
  swap_if(mask, A.y(), B.y());
  swap_if(mask, A.z(), B.z());
}
template <class Curve>
inline void swap_if_same_z(wide_mask const& mask, wide_jacobian_curve_point<Curve>& A, wide_jacobian_curve_point<Curve>& B) {
  swap_if(mask, A.x(), B.x());
  swap_if(mask, A.y(), B.y());
}
static_assert(sizeof(wide_jacobian_curve_point<>) == 384, "must match the reference's x|y|z packs");

// ---- curve_group.h ----------------------------------------------------------------------------------------
template <class Curve>
struct curve_group;
template <>
struct curve_group<curve_nist_p256> {
  using Curve = curve_nist_p256;
  using WBN = WBN256;
  using BN = bignum_256;
  using WCP = wide_curve_point<Curve>;
  using WJCP = wide_jacobian_curve_point<Curve>;

  using WMBN = ecsimd::WMBN;
  using gfp = gfp_p256;
  // curve constants in Montgomery form (curve_group.h:31-32): b*R and a*R = -3R mod p
  static BN Bm() { return bn_from_hex("dc30061d04874834e5a220abf7212ed6acf005cd78843090d89cdf6229c4bddf"); }
  static BN Am() { return bn_from_hex("fffffffc00000004000000000000000000000003fffffffffffffffffffffffc"); }
  // y with y^2 = x^3 - 3x + b (curve_group.h:43-58); empty unless all four lanes have a root
  static std::optional<gfp> compute_y(gfp const& x) {
    const gfp xpow3 = x.sqr() * x;
    const gfp x3 = gfp_shift_left<1>(x) + x;
    const gfp ypow2 = xpow3 + gfp{WMBN{WBN{Bm()}}} - x3;
    return ypow2.sqrt();
  }
  static std::optional<WBN> compute_y(WBN const& x) {
    const auto r = compute_y(gfp::from_classical(x));
    if (!r) return {};
    return {r->to_classical()};
  }
  // batch decompression: y[i] for every lane, ok[i] = 1 iff lane i has a root (npacks packs of x)
  static void compute_y(WBN* y, uint8_t* ok, WBN const* x, std::size_t npacks) { detail::check(ecb200_from_x(y, ok, x, 4 * npacks, detail::kHostPack, nullptr)); }

  static WCP WG() { return WCP{WBN{Curve::Gx()}, WBN{Curve::Gy()}}; }
  static WJCP WJG() { return WJCP::from_affine(WG()); }

  static WJCP DBLU(WJCP& P) { WJCP r, p; detail::check(ecb200_dblu(&p, &r, &P, 4, detail::kHostPack, nullptr)); P = p; return r; }
  static WJCP ZADDU(WJCP& P, WJCP const& O) { WJCP r, p; detail::check(ecb200_zaddu(&p, &r, &P, &O, 4, detail::kHostPack, nullptr)); P = p; return r; }
  static WJCP ZDAU(WJCP const& P, WJCP& Q) { WJCP r, q; detail::check(ecb200_zdau(&q, &r, &P, &Q, 4, detail::kHostPack, nullptr)); Q = q; return r; }
  static WJCP ADD_Z2_1(WJCP const& A, WJCP const& B) { WJCP r; detail::check(ecb200_add_z2_1(&r, &A, &B, 4, detail::kHostPack, nullptr)); return r; }
  static WJCP TRPLU(WJCP& P) { WJCP r, p; detail::check(ecb200_trplu(&p, &r, &P, 4, detail::kHostPack, nullptr)); P = p; return r; }

  // Scalar multiplication, 4 scalars x 4 points (curve_group.h:189-218)
  static WJCP scalar_mult(WBN const& x, WJCP P) { WJCP r; detail::check(ecb200_scalar_mult_p256(&r, &x, &P, 4, detail::kHostPack, nullptr)); return r; }
  // Scalar multiplication, 1 scalar x 4 points (curve_group.h:221-251)
  static WJCP scalar_mult_1s(BN const& x, WJCP P) {
    WJCP r;
    detail::check(ecb200_scalar_mult_p256_1s(&r, reinterpret_cast<const uint32_t*>(&x), &P, 4, detail::kHostPack, nullptr));
    return r;
  }
  // Batch: npacks packs of 4 (scalar, point) pairs in one call -- the form the GPU is built for.
  static void scalar_mult(WJCP* out, WBN const* x, WJCP const* P, std::size_t npacks) {
    detail::check(ecb200_scalar_mult_p256(out, x, P, 4 * npacks, detail::kHostPack, nullptr));
  }
  static void scalar_mult_base(WJCP* out, WBN const* x, std::size_t npacks) {
    detail::check(ecb200_scalar_mult_p256_base(out, x, 4 * npacks, detail::kHostPack, nullptr));
  }
  // scalar_mult(x, P).to_affine() for npacks packs in one call (the pair benchs/curve_group.cpp:28-46 times);
  // P == nullptr: the generator for every lane
  static void scalar_mult_affine(WCP* out, WBN const* x, WJCP const* P, std::size_t npacks) {
    detail::check(ecb200_scalar_mult_p256_affine(out, x, P, 4 * npacks, detail::kHostPack, nullptr));
  }
  static void to_affine(WCP* out, WJCP const* J, std::size_t npacks) { detail::check(ecb200_to_affine(out, J, 4 * npacks, detail::kHostPack, nullptr)); }
  static void from_affine(WJCP* out, WCP const* a, std::size_t npacks) { detail::check(ecb200_from_affine(out, a, 4 * npacks, detail::kHostPack, nullptr)); }
};

}  // namespace ecsimd

// lib/scalar_mult_p256.cpp:12-14 -- the reference's one compiled entry point
inline ecsimd::wide_jacobian_curve_point<ecsimd::curve_nist_p256> scalar_mult_p256(
    ecsimd::WBN256 const& x, ecsimd::wide_jacobian_curve_point<ecsimd::curve_nist_p256> const& P) {
  return ecsimd::curve_group<ecsimd::curve_nist_p256>::scalar_mult(x, P);
}
