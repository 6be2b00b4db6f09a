// selfcheck.cpp -- known outcomes for tests/gtest_shim/gtest/gtest.h (tests/test_reference_suite.py
// compiles this file, runs it and compares the report with the outcomes named in the test names).
#include <complex>
#include <cstdlib>
#include <limits>
#include <stdexcept>
#include <vector>

#include "gtest/gtest.h"

class Fixture : public ::testing::Test {
 protected:
  void SetUp() override { value_ = 7; }
  void TearDown() override { ++teardowns; }
  int value_ = 0;

 public:
  static int teardowns;
};
int Fixture::teardowns = 0;

class SkipInSetUp : public ::testing::Test {
 protected:
  void SetUp() override { GTEST_SKIP() << "fixture says no"; }
};

class FailInSetUp : public ::testing::Test {
 protected:
  void SetUp() override { ASSERT_EQ(1, 2); }
};

TEST_F(Fixture, Pass_SetUpRan) { ASSERT_EQ(value_, 7); }
TEST_F(Fixture, Pass_TearDownRanAfterThePreviousTest) { ASSERT_EQ(teardowns, 1); }
TEST_F(SkipInSetUp, Skip_BodyMustNotRun) { std::abort(); }
TEST_F(FailInSetUp, Fail_BodyMustNotRun) { std::abort(); }

TEST(Compare, Pass_Everything) {
  ASSERT_TRUE(2 > 1);
  ASSERT_FALSE(2 < 1);
  ASSERT_EQ(std::size_t(3), 3);
  ASSERT_EQ(std::vector<int>({1, 2}), std::vector<int>({1, 2}));
  ASSERT_EQ(std::complex<double>(1, 2), std::complex<double>(1, 2));
  ASSERT_NE(1, 2);
  ASSERT_LT(1, 2);
  ASSERT_LE(2, 2);
  ASSERT_GT(3, 2);
  ASSERT_GE(3, 3);
  ASSERT_NEAR(1.0, 1.0 + 1e-13, 1e-12);
  EXPECT_NEAR(1.0, 0.5, 0.5);
  const double one = 1.0;
  ASSERT_DOUBLE_EQ(one, one + 4 * std::numeric_limits<double>::epsilon());  // 4 units in the last place
  ASSERT_DOUBLE_EQ(0.0, -0.0);
  ASSERT_DOUBLE_EQ(-1e-320, 1e-320 - 1e-320 - 1e-320);
  if (one > 2)
    ASSERT_TRUE(false);
  else
    ASSERT_TRUE(true) << "the macros are usable in an unbraced if / else";
}
TEST(Compare, Fail_DoubleEqFiveUlps) {
  const double one = 1.0;
  ASSERT_DOUBLE_EQ(one, one + 5 * std::numeric_limits<double>::epsilon());
}
TEST(Compare, Fail_DoubleEqNan) {
  const double nan = std::numeric_limits<double>::quiet_NaN();
  ASSERT_DOUBLE_EQ(nan, nan);
}
TEST(Compare, Fail_NearNan) { ASSERT_NEAR(std::numeric_limits<double>::quiet_NaN(), 0.0, 1.0); }
TEST(Compare, Fail_NearOutside) { ASSERT_NEAR(1.0, 1.0 + 3e-12, 1e-12) << "message " << 42; }
TEST(Compare, Fail_Eq) { ASSERT_EQ(std::vector<int>({1, 2}), std::vector<int>({1, 3})); }
TEST(Compare, Fail_AssertStopsTheBody) {
  ASSERT_TRUE(false);
  std::abort();
}
TEST(Compare, Fail_ExpectGoesOn) {
  EXPECT_EQ(1, 2);
  EXPECT_LT(2, 1);
  std::cout << "after the non-fatal failures" << std::endl;
}
TEST(Skip, Skip_InBody) {
  if (true) GTEST_SKIP() << "not today";
  std::abort();
}

static void Leave(int how) {
  if (how == 0) std::abort();
  if (how == 1) std::exit(3);
  if (how == 2) throw std::runtime_error("uncaught");
}
TEST(Death, Pass_AbortExitThrow) {
  ::testing::FLAGS_gtest_death_test_style = "threadsafe";
  ASSERT_DEATH(Leave(0), "");
  EXPECT_DEATH(Leave(1), "");
  ASSERT_DEATH(Leave(2), "");
}
TEST(Death, Fail_Survives) { ASSERT_DEATH(Leave(3), ""); }

int main(int argc, char **argv) {
  ::testing::InitGoogleTest(&argc, argv);
  return RUN_ALL_TESTS();
}
