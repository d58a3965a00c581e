// oracle/integration_b200.cpp -- the binding of INTEGRATION.md section 1, built for real: this translation unit includes
// the UNMODIFIED reference headers (EVE, AVX2) next to include/ecb200.h, defines scalar_mult_p256 on top of the
// engine exactly as a maintainer of the reference would (lib/scalar_mult_p256.cpp:12-14 replaced), and then runs
//   * the reference's own scalar-mult known answers (tests/curve_group.cpp:117-173), reference types in and out;
//   * a seeded batch of packs, GPU result against curve_group<Curve>::scalar_mult of the reference on the host
//     cores, bit for bit on the Jacobian-Montgomery coordinates (and DBLU / ZADDU / ZDAU / ADD_Z2_1, mgry_mul/sqr).
// Test infrastructure: compiled on the build box by oracle/Makefile into oracle/_ref/ (the reference sources are
// not in this repository), run on the GPU box by tests/test_gpu_cpp_shim.py.
#include <ecsimd/curve_group.h>
#include <ecsimd/curve_nist_p256.h>
#include <ecsimd/literals.h>
#include <ecsimd/serialization.h>

#include <cstdio>
#include <cstring>
#include <span>
#include <stdexcept>
#include <vector>

#include <ecb200.h>

using namespace ecsimd;
using namespace ecsimd::literals;
using Curve = curve_nist_p256;
using CurveGroup = curve_group<Curve>;
using WBN = curve_wide_bn_t<Curve>;
using BN = typename WBN::value_type;
using WJCP = wide_jacobian_curve_point<Curve>;
using WMBN = curve_wide_mgry_bn_t<Curve>;
static_assert(sizeof(WBN) == 128 && sizeof(WJCP) == 384 && sizeof(WMBN) == 128);

// ---- INTEGRATION.md section 1, verbatim ---------------------------------------------------------------------------
void scalar_mult_p256_batch(std::span<WJCP> out, std::span<const WBN> x, std::span<const WJCP> P) {
  const size_t lanes = 4 * x.size();
  if (ecb200_scalar_mult_p256(out.data(), x.data(), P.data(), lanes, ECB200_LAYOUT_PACK4 | ECB200_MEM_HOST, nullptr) != ECB200_OK)
    throw std::runtime_error(ecb200_last_error());
}
auto scalar_mult_p256(WBN const& x, WJCP const& P) {
  WJCP r;
  scalar_mult_p256_batch({&r, 1}, {&x, 1}, {&P, 1});
  return r;
}
// ------------------------------------------------------------------------------------------------------------------

static int checks = 0, failures = 0;
#define CHECK(c) do { ++checks; if (!(c)) { ++failures; std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); } } while (0)
static bool same(const void* a, const void* b, size_t n) { return std::memcmp(a, b, n) == 0; }
static uint64_t splitmix(uint64_t& s) { uint64_t z = (s += 0x9e3779b97f4a7c15ull); z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull; z = (z ^ (z >> 27)) * 0x94d049bb133111ebull; return z ^ (z >> 31); }
static WBN random_wbn(uint64_t& s) {
  WBN r;
  uint64_t w[16];
  for (auto& x : w) x = splitmix(s);
  std::memcpy(&r, w, sizeof w);
  return r;
}
static void check_affine(WJCP const& J, const char* hx, const char* hy) {
  uint8_t bx[32], by[32];
  auto nib = [](char c) { return c <= '9' ? c - '0' : (c | 32) - 'a' + 10; };
  for (int i = 0; i < 32; i++) { bx[i] = uint8_t(nib(hx[2 * i]) << 4 | nib(hx[2 * i + 1])); by[i] = uint8_t(nib(hy[2 * i]) << 4 | nib(hy[2 * i + 1])); }
  const auto A = J.to_affine();   // the reference's own to_affine, on the host
  CHECK(eve::all(A.x() == WBN{bn_from_bytes_BE<BN>(bx)}));
  CHECK(eve::all(A.y() == WBN{bn_from_bytes_BE<BN>(by)}));
}

int main() {
  if (ecb200_init(0) != ECB200_OK) { std::printf("ecb200_init failed: %s\n", ecb200_last_error()); return 2; }
  const auto WJG = CurveGroup::WJG();
  {  // tests/curve_group.cpp:117-173: the three scalars, GPU result in reference types, reference to_affine
    const char* ks[3] = {"0000000000000000000000000000000000000000000000000000000000000005",
                         "0bc1b1f28709decb543d9677d2cc9942348f6b984deff409430740942ff38827",
                         "0a891cecc2bf13b0aca744434a9c9f4bd7bf5c8ed86e2f76e7df72bad813bd80"};
    const char* xs[3] = {"51590b7a515140d2d784c85608668fdfef8c82fd1f5be52421554a0dc3d033ed", "1b7721565b2c4a9f203bbccc6b531df2789fde0d135c76db71e4a7bbab9e85b2",
                         "f411d79e2997b2954975046d23b0e4a69ce580a4a81e1bed18fef6fd9ea4a912"};
    const char* ys[3] = {"e0c17da8904a727d8ae1bf36bf8a79260d012f00d4d80888d1d0bb44fda16da4", "393655bcc30f67f3a4e257b39685657d7c8df7b2a132b49c848003e300c8dcd1",
                         "43895f527937e816c3d7c0a2370002796d3cd4860cb034df86cbe7da227d9113"};
    for (int t = 0; t < 3; t++) {
      uint8_t kb[32];
      auto nib = [](char c) { return c <= '9' ? c - '0' : (c | 32) - 'a' + 10; };
      for (int i = 0; i < 32; i++) kb[i] = uint8_t(nib(ks[t][2 * i]) << 4 | nib(ks[t][2 * i + 1]));
      const WBN x{bn_from_bytes_BE<BN>(kb)};
      const WJCP gpu = scalar_mult_p256(x, WJG);
      const WJCP cpu = CurveGroup::scalar_mult(x, WJG);
      CHECK(same(&gpu, &cpu, sizeof gpu));      // the Jacobian-Montgomery representative, bit for bit
      check_affine(gpu, xs[t], ys[t]);
    }
  }
  {  // a seeded batch: 64 packs of 4 independent (scalar, point) lanes
    const size_t npacks = 64;
    uint64_t s = 0xEC51D005;
    std::vector<WBN> x(npacks);
    std::vector<WJCP> P(npacks), out(npacks);
    for (size_t i = 0; i < npacks; i++) {
      x[i] = random_wbn(s);
      P[i] = WJCP::from_affine(CurveGroup::scalar_mult(random_wbn(s), WJG).to_affine());   // r * G, Z = R
    }
    scalar_mult_p256_batch(out, x, P);
    for (size_t i = 0; i < npacks; i++) {
      const WJCP cpu = CurveGroup::scalar_mult(x[i], P[i]);
      CHECK(same(&out[i], &cpu, sizeof cpu));
    }
    // the co-Z point formulas and the field layer on the same packs, reference call by reference call
    const uint32_t F = ECB200_LAYOUT_PACK4 | ECB200_MEM_HOST;
    std::vector<WJCP> p1(npacks), d(npacks), p2(npacks), t(npacks), q(npacks), r(npacks), a(npacks);
    CHECK(ecb200_dblu(p1.data(), d.data(), P.data(), 4 * npacks, F, nullptr) == ECB200_OK);
    CHECK(ecb200_zaddu(p2.data(), t.data(), p1.data(), d.data(), 4 * npacks, F, nullptr) == ECB200_OK);
    CHECK(ecb200_zdau(q.data(), r.data(), t.data(), p2.data(), 4 * npacks, F, nullptr) == ECB200_OK);
    CHECK(ecb200_add_z2_1(a.data(), r.data(), P.data(), 4 * npacks, F, nullptr) == ECB200_OK);
    std::vector<WMBN> m(npacks), sq(npacks);
    CHECK(ecb200_mgry_mul(m.data(), &out[0].x(), &out[0].y(), 4, F, nullptr) == ECB200_OK);
    for (size_t i = 0; i < npacks; i++) {
      WJCP Pc = P[i];
      const WJCP dc = CurveGroup::DBLU(Pc);
      CHECK(same(&p1[i], &Pc, sizeof Pc) && same(&d[i], &dc, sizeof dc));
      const WJCP tc = CurveGroup::ZADDU(Pc, dc);
      CHECK(same(&p2[i], &Pc, sizeof Pc) && same(&t[i], &tc, sizeof tc));
      const WJCP rc = CurveGroup::ZDAU(tc, Pc);
      CHECK(same(&q[i], &Pc, sizeof Pc) && same(&r[i], &rc, sizeof rc));
      const WJCP ac = CurveGroup::ADD_Z2_1(rc, P[i]);
      CHECK(same(&a[i], &ac, sizeof ac));
    }
    const WMBN mc = mgry_mul(out[0].x().wmbn(), out[0].y().wmbn());
    CHECK(same(&m[0], &mc, sizeof mc));
    CHECK(ecb200_mgry_sqr(sq.data(), &out[0].x(), 4, F, nullptr) == ECB200_OK);
    const WMBN sc = mgry_sqr(out[0].x().wmbn());
    CHECK(same(&sq[0], &sc, sizeof sc));
  }
  std::printf("%s %d checks, %d failures\n", failures ? "FAILED" : "ok", checks, failures);
  return failures ? 1 : 0;
}
