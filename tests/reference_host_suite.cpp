// reference_host_suite.cpp -- the part of the reference's unit tests that never touches a register
// (conversion, TinyMatrix, ChiMatrix, Permutation, RandomNumberGenerator: unit_test/include/*_test.hpp,
// included UNCHANGED through the symlink farm oracle/_ref/dropin/unit_test/), linked against
// intel-qs_b200's libiqs.so and run on the CPU: no iqs::mpi::Environment is created, so no CUDA device
// is needed and none is used.  The full suite, with the reference's own main(), is
// oracle/_ref/dropin/bin/suite_of_tests (GPU).  Built by oracle/Makefile, run by
// tests/test_reference_suite.py.
#include <cmath>
#include <iostream>

#include "gtest/gtest.h"

#include "include/qureg.hpp"  // resolved inside the farm: -> intel-qs_b200/include/qureg.hpp

// what unit_test/suite_of_tests.cpp defines for its headers: both parts of a complex number within `error`
#define ASSERT_COMPLEX_NEAR(a, b, error) \
  ASSERT_NEAR((a).real(), (b).real(), error); \
  ASSERT_NEAR((a).imag(), (b).imag(), error);

#include "unit_test/include/conversion_test.hpp"
#include "unit_test/include/tinymatrix_test.hpp"
#include "unit_test/include/chi_matrix_test.hpp"
#include "unit_test/include/random_number_generator_test.hpp"
#include "unit_test/include/permutation_test.hpp"

int main(int argc, char **argv) {
  ::testing::InitGoogleTest(&argc, argv);
  return RUN_ALL_TESTS();
}
