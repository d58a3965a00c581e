// C++ drop-in check: the reference's own P-256 known-answer tests (tests/curve_group.cpp:38-173,
// tests/curve_point.cpp:28-42), written against include/ecsimd_b200/ecsimd.hpp with the same
// calls the reference tests make.  Built with plain g++ and linked to libecb200.so; run by
// tests/test_gpu_cpp_shim.py on the GPU box.  Prints "ok <n>" or the first failure.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../include/ecsimd_b200/ecsimd.hpp"

using namespace ecsimd;
using Curve = curve_nist_p256;
using CurveGroup = curve_group<Curve>;
using WBN = WBN256;
using WJCP = CurveGroup::WJCP;

static int checks = 0;
// comparisons are per-lane masks as in the reference: a check holds when ALL four lanes agree
#define EXPECT_TRUE(c) do { ++checks; if (!all(c)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)
static WBN set1(const char* hex) { return WBN{bn_from_hex(hex)}; }

int main() {
  if (ecb200_init(0) != 0) { std::printf("init failed: %s\n", ecb200_last_error()); return 2; }
  {  // TEST(CurveGroup, DBLU)  tests/curve_group.cpp:38-52
    auto WJG = CurveGroup::WJG();
    const auto WJdblG = CurveGroup::DBLU(WJG);
    EXPECT_TRUE(WJG.z().wbn() == WJdblG.z().wbn());
    EXPECT_TRUE(WJG.to_affine() == CurveGroup::WG());
    const auto WdblG = WJdblG.to_affine();
    EXPECT_TRUE(WdblG.x() == set1("7cf27b188d034f7e8a52380304b51ac3c08969e277f21b35a60b48fc47669978"));
    EXPECT_TRUE(WdblG.y() == set1("07775510db8ed040293d9ac69f7430dbba7dade63ce982299e04b79d227873d1"));
  }
  {  // TEST(CurveGroup, ZADDU)  :54-76   and TRPLU
    auto WJG = CurveGroup::WJG();
    const auto WJdblG = CurveGroup::DBLU(WJG);
    const auto WJ3G = CurveGroup::ZADDU(WJG, WJdblG);
    EXPECT_TRUE(WJG.z().wbn() == WJ3G.z().wbn());
    const auto W3G = WJ3G.to_affine();
    EXPECT_TRUE(W3G.x() == set1("5ecbe4d1a6330a44c8f7ef951d4bf165e6c6b721efada985fb41661bc6e7fd6c"));
    EXPECT_TRUE(W3G.y() == set1("8734640c4998ff7e374b06ce1a64a2ecd82ab036384fb83d9a79b127a27d5032"));
    auto G2 = CurveGroup::WJG();
    EXPECT_TRUE(CurveGroup::TRPLU(G2) == WJ3G);
  }
  {  // TEST(CurveGroup, ZDAU)  :78-94   2*(2G) + G = 5G
    auto WJG = CurveGroup::WJG();
    const auto WJdblG = CurveGroup::DBLU(WJG);
    const auto WJ5G = CurveGroup::ZDAU(WJdblG, WJG);
    EXPECT_TRUE(WJG.z().wbn() == WJ5G.z().wbn());
    const auto W5G = WJ5G.to_affine();
    EXPECT_TRUE(W5G.x() == set1("51590b7a515140d2d784c85608668fdfef8c82fd1f5be52421554a0dc3d033ed"));
    EXPECT_TRUE(W5G.y() == set1("e0c17da8904a727d8ae1bf36bf8a79260d012f00d4d80888d1d0bb44fda16da4"));
  }
  {  // TEST(CurveGroup, ScalarMult)  :117-173
    const auto WJG = CurveGroup::WJG();
    const auto k5 = bignum_256::from(5);
    const auto r5 = CurveGroup::scalar_mult(WBN{k5}, WJG).to_affine();
    EXPECT_TRUE(r5.x() == set1("51590b7a515140d2d784c85608668fdfef8c82fd1f5be52421554a0dc3d033ed"));
    EXPECT_TRUE(CurveGroup::scalar_mult_1s(k5, WJG).to_affine() == r5);
    const auto k = bn_from_hex("0bc1b1f28709decb543d9677d2cc9942348f6b984deff409430740942ff38827");
    const auto J = scalar_mult_p256(WBN{k}, WJG);
    // Jacobian/Montgomery representative as produced by the reference
    EXPECT_TRUE(J.x().wbn() == set1("4c315298415aa6fee7a24142ca3d3e5687e9dd69c99c308ad361c4341445835a"));
    EXPECT_TRUE(J.y().wbn() == set1("aa6cf5b34ea4ba14e76680e918bc8e19a38e60f112c49e92341052fd47611328"));
    EXPECT_TRUE(J.z().wbn() == set1("d5488a3f8e4ab4c9de98a83a0f210fed2a47ca4224eaf4f73105386f504eca20"));
    EXPECT_TRUE(CurveGroup::scalar_mult_1s(k, WJG) == J);
  }
  {  // TEST(JacobianCurvePoint, ToFromAffine)  tests/curve_point.cpp:28-42
    const auto G = CurveGroup::WG();
    EXPECT_TRUE(WJCP::from_affine(G).to_affine() == G);
  }
  {  // field ops behave like operators on GFp; batch == per-pack
    const auto a = gfp_p256::from_classical(set1("6b17d1f2e12c4247f8bce6e563a440f277037d812deb33a0f4a13945d898c296"));
    const auto b = gfp_p256::from_classical(set1("4fe342e2fe1a7f9b8ee7eb4a7c0f9e162bce33576b315ececbb6406837bf51f5"));
    EXPECT_TRUE(((a + b) - b).wbn() == a.wbn());
    EXPECT_TRUE((a * a).wbn() == a.sqr().wbn());
    EXPECT_TRUE((a * a.inverse()).wbn() == gfp_p256::one().wbn());
    EXPECT_TRUE((a + a).wbn() == gfp_shift_left<1>(a).wbn());
    EXPECT_TRUE((a + a.opposite()).to_classical() == set1("0000000000000000000000000000000000000000000000000000000000000000"));
    std::vector<WJCP> P(64, CurveGroup::WJG()), out(64);
    std::vector<WBN> ks(64);
    for (int i = 0; i < 64; i++) ks[i] = WBN{[&](int lane, int) { return bignum_256::from(uint64_t(4 * i + lane + 1)); }};
    CurveGroup::scalar_mult(out.data(), ks.data(), P.data(), 64);
    for (int i = 0; i < 64; i += 13) EXPECT_TRUE(out[i] == CurveGroup::scalar_mult(ks[i], P[i]));
    // scalar_mult(...).to_affine() as one call (benchs/curve_group.cpp:28-35), on P and on the generator
    std::vector<wide_curve_point<curve_nist_p256>> aff(64), affg(64);
    CurveGroup::scalar_mult_affine(aff.data(), ks.data(), P.data(), 64);
    CurveGroup::scalar_mult_affine(affg.data(), ks.data(), nullptr, 64);
    for (int i = 0; i < 64; i += 13) { EXPECT_TRUE(aff[i] == out[i].to_affine()); EXPECT_TRUE(affg[i] == aff[i]); }
  }
  {  // the rest of the mgry/ops + gfp + curve_point surface (mgry_ops.h:44-86, gfp.h:46-54, curve_point_ops.h:12-22,
     // curve_group.h:31-58, ifelse.h, swap.h, literals.h)
    using namespace ecsimd::literals;
    const auto x = WBN{bn_from_bytes_BE<bignum_256>("ce11d601ec0e947529e66021a0cd3d57518d58d0d5f2eb7ed75805d78c986e60"_hex)};
    const auto pt = wide_curve_point<Curve>::from_x(x);           // tests/curve_point.cpp:17-26
    EXPECT_TRUE(pt.has_value());
    EXPECT_TRUE(pt->y() == set1("f2a40cfbb248ae2c7749c76641b51b7137ccad8916931adf83b857e418fad591"));
    const auto yy = CurveGroup::compute_y(x);
    EXPECT_TRUE(yy.has_value() && all(*yy == pt->y()));
    // one lane without a root empties the optional (all four lanes or nothing)
    WBN xbad = x;
    xbad.set(2, bignum_256::from(7));   // x = 7: x^3 - 3x + b is not a square mod p
    EXPECT_TRUE(!wide_curve_point<Curve>::from_x(xbad).has_value());
    // mgry_pow: a^(p-2) is the inverse, a^((p+1)/4) squared gives a back for a square
    const auto a = gfp_p256::from_classical(x);
    const auto pm2 = bn_from_hex("ffffffff00000001000000000000000000000000fffffffffffffffffffffffd");
    EXPECT_TRUE(mgry_pow(a.wmbn(), pm2) == a.inverse().wmbn());
    const auto sq = a.sqr();
    const auto rt = sq.sqrt();
    EXPECT_TRUE(rt.has_value() && all(rt->sqr() == sq));
    // masks: if_else / swap_if / swap_if_same_z per lane
    wide_mask m = wide_mask::splat(false);
    m.set(1, true); m.set(3, true);
    auto A = CurveGroup::WJG();
    auto B2 = CurveGroup::DBLU(A);                                // A and B2 share Z
    const auto A0 = A, B0 = B2;
    const auto sel = if_else(m, A, B2);
    for (int k = 0; k < 4; k++) EXPECT_TRUE(sel.x().wbn().get(k) == (m.get(k) ? A0 : B0).x().wbn().get(k));
    swap_if_same_z(m, A, B2);
    for (int k = 0; k < 4; k++) {
      EXPECT_TRUE(A.y().wbn().get(k) == (m.get(k) ? B0 : A0).y().wbn().get(k));
      EXPECT_TRUE(B2.x().wbn().get(k) == (m.get(k) ? A0 : B0).x().wbn().get(k));
    }
    swap_if(!m, A, B2);
    swap_if(wide_mask::splat(true), A, B2);
    swap_if_same_z(m, A, B2);
    swap_if(!m, A, B2);                                            // back to (A0, B0) lane by lane
    EXPECT_TRUE(B2 == A0);
    EXPECT_TRUE(A == B0);
    // Am, Bm: the curve constants in Montgomery form
    EXPECT_TRUE(gfp_p256::from_classical(WBN{Curve::B()}).wbn() == WBN{CurveGroup::Bm()});
    EXPECT_TRUE(gfp_p256::from_classical(WBN{Curve::A()}).wbn() == WBN{CurveGroup::Am()});
  }
  std::printf("ok %d\n", checks);
  return 0;
}
