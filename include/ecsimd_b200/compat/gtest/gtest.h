// A dozen lines of GoogleTest, enough for the reference's test translation units (tests/*.cpp use TEST, EXPECT_TRUE,
// EXPECT_EQ, ASSERT_TRUE only): tests register themselves, gtest_shim::run_all() runs them and counts failures.
// GoogleTest itself is fetched from the network by the reference's CMake (CMakeLists.txt:9-17) and is not available here.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <vector>
namespace gtest_shim {
struct Case { const char* suite; const char* name; void (*fn)(); };
inline std::vector<Case>& cases() { static std::vector<Case> c; return c; }
inline int& failures() { static int f = 0; return f; }
inline int& checks() { static int n = 0; return n; }
struct Reg { Reg(const char* s, const char* n, void (*f)()) { cases().push_back({s, n, f}); } };
inline int run_all() {
  for (auto& c : cases()) {
    const int before = failures();
    c.fn();
    std::printf("[%s] %s.%s\n", failures() == before ? "  OK  " : "FAILED", c.suite, c.name);
  }
  std::printf("%s %d tests, %d checks, %d failures\n", failures() ? "FAILED" : "ok", (int)cases().size(), checks(), failures());
  return failures() ? 1 : 0;
}
}  // namespace gtest_shim
#define TEST(suite, name)                                                          \
  static void gtest_shim_##suite##_##name();                                       \
  static ::gtest_shim::Reg gtest_shim_reg_##suite##_##name(#suite, #name, &gtest_shim_##suite##_##name); \
  static void gtest_shim_##suite##_##name()
#define GTEST_SHIM_CHECK_(cond, text, fatal)                                       \
  do {                                                                             \
    ++::gtest_shim::checks();                                                      \
    if (!static_cast<bool>(cond)) {                                                \
      ++::gtest_shim::failures();                                                  \
      std::printf("%s:%d: failed: %s\n", __FILE__, __LINE__, text);               \
      if (fatal) return;                                                           \
    }                                                                              \
  } while (0)
#define EXPECT_TRUE(c) GTEST_SHIM_CHECK_((c), #c, false)
#define EXPECT_FALSE(c) GTEST_SHIM_CHECK_(!(c), "!(" #c ")", false)
#define ASSERT_TRUE(c) GTEST_SHIM_CHECK_((c), #c, true)
#define ASSERT_FALSE(c) GTEST_SHIM_CHECK_(!(c), "!(" #c ")", true)
#define EXPECT_EQ(a, b) GTEST_SHIM_CHECK_((a) == (b), #a " == " #b, false)
#define ASSERT_EQ(a, b) GTEST_SHIM_CHECK_((a) == (b), #a " == " #b, true)
#define EXPECT_NE(a, b) GTEST_SHIM_CHECK_((a) != (b), #a " != " #b, false)
