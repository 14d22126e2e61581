// tests/gtest_shim/gtest/gtest.h -- TEST INFRASTRUCTURE: a stand-in for <gtest/gtest.h> (googletest is not in this
// image) with exactly what the reference's unit tests (/root/reference/unittests/**/*.cuh) use: TEST(suite, name),
// ASSERT_* / EXPECT_* with an optional streamed message, ::testing::InitGoogleTest, RUN_ALL_TESTS().  A test that
// throws is reported as failed, not as a crash.  `--gtest_filter=a.b:c.*` (exact names or a trailing '*', ':'
// separated, a leading '-' section excludes) and `--gtest_list_tests` work as in googletest.
#pragma once

#include <cstdio>
#include <cstring>
#include <exception>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

namespace testing {

struct test_t {
  const char* suite;
  const char* name;
  void (*body)(bool&);
};
inline std::vector<test_t>& registry() {
  static std::vector<test_t> all;
  return all;
}
struct registrar_t {
  registrar_t(const char* suite, const char* name, void (*body)(bool&)) { registry().push_back({suite, name, body}); }
};
/// Collects `<< message` parts after a failed assertion and prints them when it goes out of scope.
struct sink_t {
  std::ostringstream text;
  explicit sink_t(bool) {}
  sink_t(const sink_t&) {}
  ~sink_t() {
    if (!text.str().empty())
      std::fprintf(stderr, "  %s\n", text.str().c_str());
  }
  template <typename T>
  sink_t& operator<<(const T& value) {
    text << value;
    return *this;
  }
  sink_t& operator<<(std::ostream& (*manip)(std::ostream&)) {
    text << manip;
    return *this;
  }
};
struct voidify_t {
  void operator=(const sink_t&) {}
};
inline std::string& filter() {
  static std::string f = "*";
  return f;
}
inline bool& list_only() {
  static bool l = false;
  return l;
}
inline void InitGoogleTest(int* argc, char** argv) {
  for (int i = 1; argc && i < *argc; ++i) {
    if (std::strncmp(argv[i], "--gtest_filter=", 15) == 0)
      filter() = argv[i] + 15;
    if (std::strcmp(argv[i], "--gtest_list_tests") == 0)
      list_only() = true;
  }
}
inline bool matches_one(const std::string& pattern, const std::string& full) {
  if (!pattern.empty() && pattern.back() == '*')
    return full.compare(0, pattern.size() - 1, pattern, 0, pattern.size() - 1) == 0;
  return pattern == full;
}
inline bool matches_list(const std::string& list, const std::string& full) {
  std::size_t at = 0;
  while (at <= list.size()) {
    std::size_t end = list.find(':', at);
    if (end == std::string::npos)
      end = list.size();
    if (end > at && matches_one(list.substr(at, end - at), full))
      return true;
    at = end + 1;
  }
  return false;
}
inline bool selected(const std::string& full) {
  const std::string& f = filter();
  const std::size_t dash = f.find('-');
  const std::string positive = dash == std::string::npos ? f : f.substr(0, dash);
  const std::string negative = dash == std::string::npos ? "" : f.substr(dash + 1);
  return (positive.empty() || matches_list(positive, full)) && !(negative.size() && matches_list(negative, full));
}

}  // namespace testing

#define TEST(suite, name)                                                                        \
  static void gtest_##suite##_##name##_body(bool& gtest_failed);                                 \
  static ::testing::registrar_t gtest_##suite##_##name##_registrar(#suite, #name,                \
                                                                   gtest_##suite##_##name##_body); \
  static void gtest_##suite##_##name##_body(bool& gtest_failed)

#define GTEST_SHIM_FAIL_(text) \
  ::testing::sink_t((gtest_failed = true, std::fprintf(stderr, "%s:%d: Failure: %s\n", __FILE__, __LINE__, text), true))
#define GTEST_SHIM_ASSERT_(condition, text) \
  if (condition) {                          \
  } else                                    \
    return ::testing::voidify_t() = GTEST_SHIM_FAIL_(text)
#define GTEST_SHIM_EXPECT_(condition, text) \
  if (condition) {                          \
  } else                                    \
    ::testing::voidify_t() = GTEST_SHIM_FAIL_(text)

#define ASSERT_EQ(a, b) GTEST_SHIM_ASSERT_((a) == (b), "ASSERT_EQ(" #a ", " #b ")")
#define ASSERT_NE(a, b) GTEST_SHIM_ASSERT_((a) != (b), "ASSERT_NE(" #a ", " #b ")")
#define ASSERT_LT(a, b) GTEST_SHIM_ASSERT_((a) < (b), "ASSERT_LT(" #a ", " #b ")")
#define ASSERT_LE(a, b) GTEST_SHIM_ASSERT_((a) <= (b), "ASSERT_LE(" #a ", " #b ")")
#define ASSERT_GT(a, b) GTEST_SHIM_ASSERT_((a) > (b), "ASSERT_GT(" #a ", " #b ")")
#define ASSERT_GE(a, b) GTEST_SHIM_ASSERT_((a) >= (b), "ASSERT_GE(" #a ", " #b ")")
#define ASSERT_TRUE(a) GTEST_SHIM_ASSERT_(static_cast<bool>(a), "ASSERT_TRUE(" #a ")")
#define ASSERT_FALSE(a) GTEST_SHIM_ASSERT_(!static_cast<bool>(a), "ASSERT_FALSE(" #a ")")
#define EXPECT_EQ(a, b) GTEST_SHIM_EXPECT_((a) == (b), "EXPECT_EQ(" #a ", " #b ")")
#define EXPECT_NE(a, b) GTEST_SHIM_EXPECT_((a) != (b), "EXPECT_NE(" #a ", " #b ")")
#define EXPECT_LT(a, b) GTEST_SHIM_EXPECT_((a) < (b), "EXPECT_LT(" #a ", " #b ")")
#define EXPECT_LE(a, b) GTEST_SHIM_EXPECT_((a) <= (b), "EXPECT_LE(" #a ", " #b ")")
#define EXPECT_GT(a, b) GTEST_SHIM_EXPECT_((a) > (b), "EXPECT_GT(" #a ", " #b ")")
#define EXPECT_GE(a, b) GTEST_SHIM_EXPECT_((a) >= (b), "EXPECT_GE(" #a ", " #b ")")
#define EXPECT_TRUE(a) GTEST_SHIM_EXPECT_(static_cast<bool>(a), "EXPECT_TRUE(" #a ")")
#define EXPECT_FALSE(a) GTEST_SHIM_EXPECT_(!static_cast<bool>(a), "EXPECT_FALSE(" #a ")")

inline int RUN_ALL_TESTS() {
  int ran = 0, bad = 0;
  for (auto& t : ::testing::registry()) {
    const std::string full = std::string(t.suite) + "." + t.name;
    if (!::testing::selected(full))
      continue;
    if (::testing::list_only()) {
      std::printf("%s\n", full.c_str());
      continue;
    }
    std::printf("[ RUN      ] %s\n", full.c_str());
    std::fflush(stdout);
    bool failed = false;
    try {
      t.body(failed);
    } catch (const std::exception& e) {
      std::fprintf(stderr, "  exception: %s\n", e.what());
      failed = true;
    } catch (...) {
      std::fprintf(stderr, "  unknown exception\n");
      failed = true;
    }
    std::printf(failed ? "[  FAILED  ] %s\n" : "[       OK ] %s\n", full.c_str());
    ++ran;
    bad += failed;
  }
  if (!::testing::list_only())
    std::printf("[==========] %d tests ran, %d failed\n", ran, bad);
  return bad ? 1 : 0;
}
