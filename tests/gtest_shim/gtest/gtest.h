// gtest/gtest.h -- a small stand-in for googletest, test infrastructure.
//
// The reference's own unit tests (unit_test/suite_of_tests.cpp + unit_test/include/*_test.hpp) are
// written against googletest, which its CMake fetches from the network; this image has neither.  This
// header implements the part of googletest's public macro interface those files use, so that the suite
// compiles UNCHANGED (oracle/Makefile: _ref/refbin/suite_of_tests against the reference library -- that
// run validates this header on the CPU -- and _ref/dropin/bin/suite_of_tests against intel-qs_b200):
//
//   TEST, TEST_F, ::testing::Test (SetUp / TearDown), GTEST_SKIP,
//   ASSERT_/EXPECT_ {TRUE, FALSE, EQ, NE, LT, LE, GT, GE, NEAR, DOUBLE_EQ, FLOAT_EQ, DEATH}, `<< message`,
//   ::testing::InitGoogleTest (--gtest_filter=, --gtest_list_tests), RUN_ALL_TESTS,
//   ::testing::UnitTest::GetInstance()->listeners() (Release / default_result_printer),
//   ::testing::FLAGS_gtest_death_test_style.
//
// The output lines ([ RUN      ], [       OK ], [  SKIPPED ], [  FAILED  ], [  PASSED  ] N tests.) follow
// googletest's so that logs read the same.  Semantics that matter to the suite:
//   * a fatal failure or a skip inside SetUp() keeps the body from running (TearDown still runs);
//   * ASSERT_DOUBLE_EQ accepts a distance of at most 4 units in the last place, NaN never compares equal;
//   * a death test re-executes the test binary (googletest's "threadsafe" style: a forked copy of a process
//     with OpenMP worker threads or a CUDA context cannot run the statement) restricted to the current
//     test; the child runs the test up to the death statement, executes it with its output silenced and
//     leaves with status 0 if it survived.  The parent passes when the child ended by a signal or a
//     non-zero status.
#ifndef IQS_B200_GTEST_SHIM_H
#define IQS_B200_GTEST_SHIM_H

#include <sys/types.h>
#include <sys/wait.h>
#include <unistd.h>

#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <iomanip>
#include <iostream>
#include <iterator>
#include <limits>
#include <memory>
#include <sstream>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

namespace testing {

// ---------------------------------------------------------------- printing of compared values
namespace internal {

template <class...>
using void_t = void;

template <class T, class = void>
struct is_streamable : std::false_type {};
template <class T>
struct is_streamable<T, void_t<decltype(std::declval<std::ostream &>() << std::declval<const T &>())>> : std::true_type {};

template <class T, class = void>
struct is_iterable : std::false_type {};
template <class T>
struct is_iterable<T, void_t<decltype(std::begin(std::declval<const T &>())), decltype(std::end(std::declval<const T &>()))>> : std::true_type {};

template <class T>
void PrintValue(std::ostream &os, const T &v);

template <class T>
typename std::enable_if<is_streamable<T>::value>::type PrintImpl(std::ostream &os, const T &v) {
  os << v;
}
template <class T>
typename std::enable_if<!is_streamable<T>::value && is_iterable<T>::value>::type PrintImpl(std::ostream &os, const T &v) {
  os << "{";
  bool first = true;
  for (const auto &e : v) {
    if (!first) os << ", ";
    first = false;
    PrintValue(os, e);
  }
  os << "}";
}
template <class T>
typename std::enable_if<!is_streamable<T>::value && !is_iterable<T>::value>::type PrintImpl(std::ostream &os, const T &) {
  os << "<" << sizeof(T) << "-byte object>";
}

template <class T>
void PrintValue(std::ostream &os, const T &v) {
  PrintImpl(os, v);
}
inline void PrintValue(std::ostream &os, bool v) { os << (v ? "true" : "false"); }
inline void PrintValue(std::ostream &os, char v) { os << "'" << v << "' (" << int(v) << ")"; }
inline void PrintValue(std::ostream &os, signed char v) { os << int(v); }
inline void PrintValue(std::ostream &os, unsigned char v) { os << unsigned(v); }
inline void PrintValue(std::ostream &os, float v) { os << std::setprecision(9) << v; }
inline void PrintValue(std::ostream &os, double v) { os << std::setprecision(17) << v; }
inline void PrintValue(std::ostream &os, std::nullptr_t) { os << "nullptr"; }
inline void PrintValue(std::ostream &os, const std::string &v) { os << '"' << v << '"'; }
inline void PrintValue(std::ostream &os, const char *v) {
  if (v) os << '"' << v << '"';
  else os << "NULL";
}

template <class T>
std::string ToString(const T &v) {
  std::ostringstream os;
  PrintValue(os, v);
  return os.str();
}

}  // namespace internal

// ---------------------------------------------------------------- user message: ASSERT_X(...) << "text" << value
class Message {
 public:
  Message() {}
  Message(const Message &o) { ss_ << o.str(); }
  template <class T>
  Message &operator<<(const T &v) {
    ss_ << v;
    return *this;
  }
  Message &operator<<(std::ostream &(*manip)(std::ostream &)) {
    ss_ << manip;
    return *this;
  }
  Message &operator<<(bool b) {
    ss_ << (b ? "true" : "false");
    return *this;
  }
  std::string str() const { return ss_.str(); }

 private:
  std::ostringstream ss_;
};

// ---------------------------------------------------------------- outcome of one comparison
class AssertionResult {
 public:
  explicit AssertionResult(bool ok) : ok_(ok) {}
  AssertionResult(bool ok, std::string text) : ok_(ok), text_(std::move(text)) {}
  explicit operator bool() const { return ok_; }
  const std::string &text() const { return text_; }

 private:
  bool ok_;
  std::string text_;
};

inline AssertionResult AssertionSuccess() { return AssertionResult(true); }
inline AssertionResult AssertionFailure(const std::string &text = std::string()) { return AssertionResult(false, text); }

// ---------------------------------------------------------------- the registry and the state of the running test
class Test;

namespace internal {

struct TestInfo {
  std::string suite, name;
  std::function<Test *()> make;
};

struct State {
  std::vector<TestInfo> tests;
  std::string filter = "*";
  bool list_only = false;
  bool printing = true;  // the default result printer is attached
  std::string death_site;  // "file:line" of the one death statement this (child) process is to execute
  // per running test
  std::string current;
  bool failed = false, fatal = false, skipped = false;
};

inline State &S() {
  static State s;
  return s;
}

inline int Register(const char *suite, const char *name, std::function<Test *()> make) {
  S().tests.push_back(TestInfo{suite, name, std::move(make)});
  return 0;
}

enum class Kind { kNonFatal, kFatal, kSkip };

// `return AssertHelper(...) = Message() << ...;` -- operator= returns void so that the statement is usable
// in void functions, like googletest's own helper.
class AssertHelper {
 public:
  AssertHelper(Kind kind, const char *file, int line, std::string text) : kind_(kind), file_(file), line_(line), text_(std::move(text)) {}
  void operator=(const Message &m) const {
    State &s = S();
    const std::string user = m.str();
    if (kind_ == Kind::kSkip) {
      s.skipped = true;
      if (s.printing) {
        std::cout << file_ << ":" << line_ << ": Skipped" << std::endl;
        if (!user.empty()) std::cout << user << std::endl;
      }
      return;
    }
    s.failed = true;
    if (kind_ == Kind::kFatal) s.fatal = true;
    if (s.printing) {
      std::cout << file_ << ":" << line_ << ": Failure" << std::endl << text_ << std::endl;
      if (!user.empty()) std::cout << user << std::endl;
    }
  }

 private:
  Kind kind_;
  const char *file_;
  int line_;
  std::string text_;
};

// glob with '*' and '?', patterns separated by ':', negative part after '-'
inline bool GlobMatch(const char *p, const char *s) {
  if (*p == 0) return *s == 0;
  if (*p == '*') return GlobMatch(p + 1, s) || (*s != 0 && GlobMatch(p, s + 1));
  if (*s == 0) return false;
  return (*p == '?' || *p == *s) && GlobMatch(p + 1, s + 1);
}
inline bool AnyGlob(const std::string &patterns, const std::string &name) {
  std::size_t start = 0;
  while (start <= patterns.size()) {
    std::size_t end = patterns.find(':', start);
    if (end == std::string::npos) end = patterns.size();
    const std::string one = patterns.substr(start, end - start);
    if (!one.empty() && GlobMatch(one.c_str(), name.c_str())) return true;
    start = end + 1;
  }
  return false;
}
inline bool FilterAccepts(const std::string &filter, const std::string &full) {
  const std::size_t dash = filter.find('-');
  const std::string pos = dash == std::string::npos ? filter : filter.substr(0, dash);
  const std::string neg = dash == std::string::npos ? std::string() : filter.substr(dash + 1);
  return AnyGlob(pos.empty() ? std::string("*") : pos, full) && !AnyGlob(neg, full);
}

// ---- comparisons -------------------------------------------------------------------------------
template <class A, class B>
std::string CmpText(const char *ea, const char *eb, const A &a, const B &b, const char *op) {
  std::ostringstream os;
  os << "Expected: (" << ea << ") " << op << " (" << eb << "), actual: " << ToString(a) << " vs " << ToString(b);
  return os.str();
}

#define IQS_GTEST_SHIM_CMP(Name, op)                                                      \
  template <class A, class B>                                                             \
  AssertionResult Cmp##Name(const char *ea, const char *eb, const A &a, const B &b) {     \
    if (a op b) return AssertionSuccess();                                                \
    return AssertionFailure(CmpText(ea, eb, a, b, #op));                                  \
  }
IQS_GTEST_SHIM_CMP(NE, !=)
IQS_GTEST_SHIM_CMP(LT, <)
IQS_GTEST_SHIM_CMP(LE, <=)
IQS_GTEST_SHIM_CMP(GT, >)
IQS_GTEST_SHIM_CMP(GE, >=)
#undef IQS_GTEST_SHIM_CMP

template <class A, class B>
AssertionResult CmpEQ(const char *ea, const char *eb, const A &a, const B &b) {
  if (a == b) return AssertionSuccess();
  std::ostringstream os;
  os << "Expected equality of these values:\n  " << ea << "\n    Which is: " << ToString(a) << "\n  " << eb << "\n    Which is: " << ToString(b);
  return AssertionFailure(os.str());
}

// distance in units in the last place via the biased integer representation (sign-magnitude -> offset)
template <class F, class U>
bool AlmostEqualUlps(F a, F b) {
  if (std::isnan(a) || std::isnan(b)) return false;
  U ua, ub;
  std::memcpy(&ua, &a, sizeof(F));
  std::memcpy(&ub, &b, sizeof(F));
  const U sign = U(1) << (8 * sizeof(U) - 1);
  const U ba = (ua & sign) ? U(~ua + 1) : U(sign | ua);
  const U bb = (ub & sign) ? U(~ub + 1) : U(sign | ub);
  const U dist = ba >= bb ? ba - bb : bb - ba;
  return dist <= 4;
}

inline AssertionResult CmpDoubleEQ(const char *ea, const char *eb, double a, double b) {
  if (AlmostEqualUlps<double, std::uint64_t>(a, b)) return AssertionSuccess();
  std::ostringstream os;
  os << "Expected equality of these values:\n  " << ea << "\n    Which is: " << ToString(a) << "\n  " << eb << "\n    Which is: " << ToString(b);
  return AssertionFailure(os.str());
}
inline AssertionResult CmpFloatEQ(const char *ea, const char *eb, float a, float b) {
  if (AlmostEqualUlps<float, std::uint32_t>(a, b)) return AssertionSuccess();
  std::ostringstream os;
  os << "Expected equality of these values:\n  " << ea << "\n    Which is: " << ToString(a) << "\n  " << eb << "\n    Which is: " << ToString(b);
  return AssertionFailure(os.str());
}
inline AssertionResult CmpNear(const char *ea, const char *eb, const char *ee, double a, double b, double err) {
  const double diff = std::fabs(a - b);
  if (diff <= err) return AssertionSuccess();  // false for NaN, as it should be
  std::ostringstream os;
  os << "The difference between " << ea << " and " << eb << " is " << ToString(diff) << ", which exceeds " << ee << ", where\n"
     << ea << " evaluates to " << ToString(a) << ",\n" << eb << " evaluates to " << ToString(b) << ", and\n" << ee << " evaluates to " << ToString(err) << ".";
  return AssertionFailure(os.str());
}
inline AssertionResult CmpBool(const char *expr, bool value, bool expected) {
  if (value == expected) return AssertionSuccess();
  std::ostringstream os;
  os << "Value of: " << expr << "\n  Actual: " << (value ? "true" : "false") << "\nExpected: " << (expected ? "true" : "false");
  return AssertionFailure(os.str());
}

// ---- death tests -------------------------------------------------------------------------------
// Everything the suite kills itself with (assert, abort, an uncaught exception, exit(1)) ends the child
// abnormally; surviving the statement ends it with status 0.
template <class Fn>
AssertionResult Dies(const char *stmt, const char *file, int line, Fn &&fn) {
  State &s = S();
  const std::string site = std::string(file) + ":" + std::to_string(line);
  if (!s.death_site.empty()) {  // this process IS a death-test child
    if (s.death_site != site) return AssertionSuccess();  // some other death statement of the same test
    std::fflush(nullptr);
    if (FILE *nul = std::freopen("/dev/null", "w", stderr)) (void)nul;
    if (FILE *nul = std::freopen("/dev/null", "w", stdout)) (void)nul;
    fn();
    _exit(0);
  }
  std::cout.flush();
  std::fflush(nullptr);
  const std::string filter = "--gtest_filter=" + s.current, flag = "--gtest_internal_run_death_test=" + site;
  const pid_t pid = fork();
  if (pid < 0) return AssertionFailure(std::string("fork failed for death test: ") + stmt);
  if (pid == 0) {  // nothing but exec between fork and the new image
    execl("/proc/self/exe", "death_test_child", filter.c_str(), flag.c_str(), (char *)nullptr);
    _exit(0);  // exec failed: "survived", the parent reports it
  }
  int status = 0;
  while (waitpid(pid, &status, 0) < 0) {
  }
  if (WIFSIGNALED(status) || (WIFEXITED(status) && WEXITSTATUS(status) != 0)) return AssertionSuccess();
  return AssertionFailure(std::string("Death test: ") + stmt + "\n    Result: failed to die.");
}

}  // namespace internal

// ---------------------------------------------------------------- base class of the fixtures
class Test {
 public:
  virtual ~Test() {}
  static bool HasFatalFailure() { return internal::S().fatal; }
  static bool HasNonfatalFailure() { return internal::S().failed && !internal::S().fatal; }
  static bool HasFailure() { return internal::S().failed; }
  static bool IsSkipped() { return internal::S().skipped; }

  // run by RUN_ALL_TESTS
  void Run() {
    SetUp();
    if (!HasFatalFailure() && !IsSkipped()) TestBody();
    TearDown();
  }

 protected:
  Test() {}
  virtual void SetUp() {}
  virtual void TearDown() {}
  virtual void TestBody() = 0;
};

// ---------------------------------------------------------------- the little of the listener API the suite's main() touches
class TestEventListener {
 public:
  virtual ~TestEventListener() {}
};

class TestEventListeners {
 public:
  TestEventListener *default_result_printer() const { return printer_.get(); }
  // detaches the listener and hands its ownership to the caller
  TestEventListener *Release(TestEventListener *listener) {
    if (listener != nullptr && listener == printer_.get()) {
      internal::S().printing = false;
      return printer_.release();
    }
    return nullptr;
  }

 private:
  std::unique_ptr<TestEventListener> printer_{new TestEventListener};
};

class UnitTest {
 public:
  static UnitTest *GetInstance() {
    static UnitTest instance;
    return &instance;
  }
  TestEventListeners &listeners() { return listeners_; }

  int Run() {
    internal::State &s = internal::S();
    std::vector<const internal::TestInfo *> chosen;
    for (const auto &t : s.tests)
      if (internal::FilterAccepts(s.filter, t.suite + "." + t.name)) chosen.push_back(&t);
    if (s.list_only) {
      std::string last;
      for (const auto *t : chosen) {
        if (t->suite != last) std::cout << t->suite << "." << std::endl;
        last = t->suite;
        std::cout << "  " << t->name << std::endl;
      }
      return 0;
    }
    if (!s.death_site.empty()) s.printing = false;  // a death-test child reports through its exit status only
    const bool out = s.printing;
    if (out) std::cout << "[==========] Running " << chosen.size() << " tests." << std::endl;
    std::vector<std::string> failed, skipped;
    std::size_t passed = 0;
    for (const auto *t : chosen) {
      const std::string full = t->suite + "." + t->name;
      s.failed = s.fatal = s.skipped = false;
      s.current = full;
      if (out) std::cout << "[ RUN      ] " << full << std::endl;
      {
        std::unique_ptr<Test> test(t->make());
        test->Run();
      }
      if (s.failed) {
        failed.push_back(full);
        if (out) std::cout << "[  FAILED  ] " << full << std::endl;
      } else if (s.skipped) {
        skipped.push_back(full);
        if (out) std::cout << "[  SKIPPED ] " << full << std::endl;
      } else {
        ++passed;
        if (out) std::cout << "[       OK ] " << full << std::endl;
      }
    }
    if (out) {
      std::cout << "[==========] " << chosen.size() << " tests ran." << std::endl;
      std::cout << "[  PASSED  ] " << passed << " tests." << std::endl;
      if (!skipped.empty()) {
        std::cout << "[  SKIPPED ] " << skipped.size() << " tests, listed below:" << std::endl;
        for (const auto &n : skipped) std::cout << "[  SKIPPED ] " << n << std::endl;
      }
      if (!failed.empty()) {
        std::cout << "[  FAILED  ] " << failed.size() << " tests, listed below:" << std::endl;
        for (const auto &n : failed) std::cout << "[  FAILED  ] " << n << std::endl;
      }
    }
    return failed.empty() ? 0 : 1;
  }

 private:
  UnitTest() {}
  TestEventListeners listeners_;
};

// The flag the suite assigns to (the value is ignored: every death test here is "threadsafe").
// `inline` variables are C++17; a function-local static behind a reference keeps this header C++14.
inline std::string &DeathTestStyleFlag() {
  static std::string style = "fast";
  return style;
}
static std::string &FLAGS_gtest_death_test_style = DeathTestStyleFlag();

inline void InitGoogleTest(int *argc, char **argv) {
  internal::State &s = internal::S();
  int kept = 1;
  for (int i = 1; argc != nullptr && i < *argc; ++i) {
    const std::string a = argv[i];
    if (a.rfind("--gtest_filter=", 0) == 0) s.filter = a.substr(15);
    else if (a == "--gtest_list_tests") s.list_only = true;
    else if (a.rfind("--gtest_internal_run_death_test=", 0) == 0) s.death_site = a.substr(32);
    else if (a.rfind("--gtest_", 0) == 0) {
    }  // other googletest flags: accepted, no effect
    else argv[kept++] = argv[i];
  }
  if (argc != nullptr && *argc > 0) *argc = kept;
  if (const char *env = std::getenv("GTEST_FILTER"))
    if (s.filter == "*") s.filter = env;
}
inline void InitGoogleTest() {}

}  // namespace testing

inline int RUN_ALL_TESTS() { return ::testing::UnitTest::GetInstance()->Run(); }

// ---------------------------------------------------------------- macros
#define IQS_GTEST_SHIM_CLASS(suite, name) suite##_##name##_Test

#define IQS_GTEST_SHIM_TEST(suite, name, parent)                                                                  \
  class IQS_GTEST_SHIM_CLASS(suite, name) : public parent {                                                       \
   public:                                                                                                        \
    IQS_GTEST_SHIM_CLASS(suite, name)() {}                                                                        \
                                                                                                                  \
   private:                                                                                                       \
    void TestBody() override;                                                                                     \
    static int registered_;                                                                                       \
  };                                                                                                              \
  int IQS_GTEST_SHIM_CLASS(suite, name)::registered_ = ::testing::internal::Register(                             \
      #suite, #name, []() -> ::testing::Test * { return new IQS_GTEST_SHIM_CLASS(suite, name); });                \
  void IQS_GTEST_SHIM_CLASS(suite, name)::TestBody()

#define TEST(suite, name) IQS_GTEST_SHIM_TEST(suite, name, ::testing::Test)
#define TEST_F(fixture, name) IQS_GTEST_SHIM_TEST(fixture, name, fixture)

// `if (result) ; else <report>` lets a trailing `<< message` bind to the report; the dangling-else form is
// the same one googletest uses, so the macros behave alike inside unbraced if/else.
#define IQS_GTEST_SHIM_CHECK(result_expr, kind, on_fail)                                        \
  switch (0)                                                                                    \
  case 0:                                                                                       \
  default:                                                                                      \
    if (const ::testing::AssertionResult iqs_gtest_ar = (result_expr))                          \
      ;                                                                                         \
    else                                                                                        \
      on_fail ::testing::internal::AssertHelper(kind, __FILE__, __LINE__, iqs_gtest_ar.text()) = ::testing::Message()

#define IQS_GTEST_SHIM_FATAL(result_expr) IQS_GTEST_SHIM_CHECK(result_expr, ::testing::internal::Kind::kFatal, return)
#define IQS_GTEST_SHIM_NONFATAL(result_expr) IQS_GTEST_SHIM_CHECK(result_expr, ::testing::internal::Kind::kNonFatal, )

#define ASSERT_TRUE(c) IQS_GTEST_SHIM_FATAL(::testing::internal::CmpBool(#c, static_cast<bool>(c), true))
#define ASSERT_FALSE(c) IQS_GTEST_SHIM_FATAL(::testing::internal::CmpBool(#c, static_cast<bool>(c), false))
#define EXPECT_TRUE(c) IQS_GTEST_SHIM_NONFATAL(::testing::internal::CmpBool(#c, static_cast<bool>(c), true))
#define EXPECT_FALSE(c) IQS_GTEST_SHIM_NONFATAL(::testing::internal::CmpBool(#c, static_cast<bool>(c), false))

#define ASSERT_EQ(a, b) IQS_GTEST_SHIM_FATAL(::testing::internal::CmpEQ(#a, #b, a, b))
#define ASSERT_NE(a, b) IQS_GTEST_SHIM_FATAL(::testing::internal::CmpNE(#a, #b, a, b))
#define ASSERT_LT(a, b) IQS_GTEST_SHIM_FATAL(::testing::internal::CmpLT(#a, #b, a, b))
#define ASSERT_LE(a, b) IQS_GTEST_SHIM_FATAL(::testing::internal::CmpLE(#a, #b, a, b))
#define ASSERT_GT(a, b) IQS_GTEST_SHIM_FATAL(::testing::internal::CmpGT(#a, #b, a, b))
#define ASSERT_GE(a, b) IQS_GTEST_SHIM_FATAL(::testing::internal::CmpGE(#a, #b, a, b))
#define EXPECT_EQ(a, b) IQS_GTEST_SHIM_NONFATAL(::testing::internal::CmpEQ(#a, #b, a, b))
#define EXPECT_NE(a, b) IQS_GTEST_SHIM_NONFATAL(::testing::internal::CmpNE(#a, #b, a, b))
#define EXPECT_LT(a, b) IQS_GTEST_SHIM_NONFATAL(::testing::internal::CmpLT(#a, #b, a, b))
#define EXPECT_LE(a, b) IQS_GTEST_SHIM_NONFATAL(::testing::internal::CmpLE(#a, #b, a, b))
#define EXPECT_GT(a, b) IQS_GTEST_SHIM_NONFATAL(::testing::internal::CmpGT(#a, #b, a, b))
#define EXPECT_GE(a, b) IQS_GTEST_SHIM_NONFATAL(::testing::internal::CmpGE(#a, #b, a, b))

#define ASSERT_DOUBLE_EQ(a, b) IQS_GTEST_SHIM_FATAL(::testing::internal::CmpDoubleEQ(#a, #b, a, b))
#define EXPECT_DOUBLE_EQ(a, b) IQS_GTEST_SHIM_NONFATAL(::testing::internal::CmpDoubleEQ(#a, #b, a, b))
#define ASSERT_FLOAT_EQ(a, b) IQS_GTEST_SHIM_FATAL(::testing::internal::CmpFloatEQ(#a, #b, a, b))
#define EXPECT_FLOAT_EQ(a, b) IQS_GTEST_SHIM_NONFATAL(::testing::internal::CmpFloatEQ(#a, #b, a, b))
#define ASSERT_NEAR(a, b, e) IQS_GTEST_SHIM_FATAL(::testing::internal::CmpNear(#a, #b, #e, a, b, e))
#define EXPECT_NEAR(a, b, e) IQS_GTEST_SHIM_NONFATAL(::testing::internal::CmpNear(#a, #b, #e, a, b, e))

#define ASSERT_DEATH(stmt, regex) IQS_GTEST_SHIM_FATAL(::testing::internal::Dies(#stmt, __FILE__, __LINE__, [&]() { stmt; }))
#define EXPECT_DEATH(stmt, regex) IQS_GTEST_SHIM_NONFATAL(::testing::internal::Dies(#stmt, __FILE__, __LINE__, [&]() { stmt; }))

#define GTEST_SKIP() \
  return ::testing::internal::AssertHelper(::testing::internal::Kind::kSkip, __FILE__, __LINE__, std::string()) = ::testing::Message()
#define GTEST_FAIL() \
  return ::testing::internal::AssertHelper(::testing::internal::Kind::kFatal, __FILE__, __LINE__, "Failed") = ::testing::Message()
#define FAIL() GTEST_FAIL()
#define ADD_FAILURE() \
  ::testing::internal::AssertHelper(::testing::internal::Kind::kNonFatal, __FILE__, __LINE__, "Failed") = ::testing::Message()
#define SUCCEED() ::testing::Message()

#endif  // IQS_B200_GTEST_SHIM_H
