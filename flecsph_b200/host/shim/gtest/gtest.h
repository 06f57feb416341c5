// Minimal stand-in for the part of GoogleTest the reference's unit tests use (mpisph/test/bs.cc, io.cc: TEST, ASSERT_TRUE,
// ...), with a main() that runs the registered tests -- so that those test sources build UNMODIFIED here (no GTest in
// the image), against the reference's own classes (oracle/_ref/*_test_ref) and against the GPU facade
// (flecsph_b200/host/bin/*_test_b200).  A failed assertion prints file:line and the expression, marks the test failed and
// returns from the test body, as ASSERT_* does.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <utility>
#include <vector>

namespace fsph_gtest {
typedef void (*body_t)();
inline std::vector<std::pair<const char *, body_t>> & all() {
  static std::vector<std::pair<const char *, body_t>> v;
  return v;
}
inline int & failed() {
  static int f = 0;
  return f;
}
struct reg {
  reg(const char * name, body_t fn) { all().push_back({name, fn}); }
};
}  // namespace fsph_gtest

#define TEST(suite, name)                                                             \
  static void suite##_##name##_body();                                                \
  static fsph_gtest::reg suite##_##name##_reg(#suite "." #name, suite##_##name##_body); \
  static void suite##_##name##_body()

#define FSPH_GTEST_CHECK(cond, text, fatal)                                          \
  do {                                                                                \
    if(!(cond)) {                                                                     \
      std::fprintf(stderr, "%s:%d: Failure: %s\n", __FILE__, __LINE__, text);         \
      ++fsph_gtest::failed();                                                         \
      if(fatal) return;                                                               \
    }                                                                                 \
  } while(0)
#define ASSERT_TRUE(x) FSPH_GTEST_CHECK((x), "ASSERT_TRUE(" #x ")", true)
#define ASSERT_FALSE(x) FSPH_GTEST_CHECK(!(x), "ASSERT_FALSE(" #x ")", true)
#define ASSERT_EQ(a, b) FSPH_GTEST_CHECK((a) == (b), "ASSERT_EQ(" #a ", " #b ")", true)
#define ASSERT_NE(a, b) FSPH_GTEST_CHECK((a) != (b), "ASSERT_NE(" #a ", " #b ")", true)
#define ASSERT_LT(a, b) FSPH_GTEST_CHECK((a) < (b), "ASSERT_LT(" #a ", " #b ")", true)
#define ASSERT_LE(a, b) FSPH_GTEST_CHECK((a) <= (b), "ASSERT_LE(" #a ", " #b ")", true)
#define ASSERT_GT(a, b) FSPH_GTEST_CHECK((a) > (b), "ASSERT_GT(" #a ", " #b ")", true)
#define ASSERT_GE(a, b) FSPH_GTEST_CHECK((a) >= (b), "ASSERT_GE(" #a ", " #b ")", true)
#define EXPECT_TRUE(x) FSPH_GTEST_CHECK((x), "EXPECT_TRUE(" #x ")", false)
#define EXPECT_FALSE(x) FSPH_GTEST_CHECK(!(x), "EXPECT_FALSE(" #x ")", false)
#define EXPECT_EQ(a, b) FSPH_GTEST_CHECK((a) == (b), "EXPECT_EQ(" #a ", " #b ")", false)

int main(int, char **) {
  int bad = 0;
  for(auto & t : fsph_gtest::all()) {
    std::printf("[ RUN      ] %s\n", t.first);
    const int before = fsph_gtest::failed();
    t.second();
    const bool ok = fsph_gtest::failed() == before;
    std::printf("[ %s ] %s\n", ok ? "      OK" : " FAILED ", t.first);
    bad += ok ? 0 : 1;
  }
  std::printf("[==========] %zu tests ran, %d failed\n", fsph_gtest::all().size(), bad);
  return bad ? 1 : 0;
}
