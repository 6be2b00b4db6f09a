// host_logic.cpp -- exercises the host-only classes of the Intel-QS API through their public
// interface and prints the results.  TEST INFRASTRUCTURE: compiled once against the reference's
// headers/library and once against intel-qs_b200/include + libiqs.so; the two outputs must be
// identical (tests/test_host_logic.py).  No state vector is created, so no GPU is needed.
#include <cstdio>
#include <random>
#include <vector>

#include "qureg.hpp"

static void print_vec(const char *tag, const std::vector<std::size_t> &v) {
  printf("%s", tag);
  for (auto x : v) printf(" %zu", x);
  printf("\n");
}

int main() {
  // --- Permutation ---------------------------------------------------------------------------
  std::mt19937 gen(12345);
  for (int trial = 0; trial < 40; ++trial) {
    std::size_t n = 3 + trial % 9;
    std::vector<std::size_t> a(n), b(n);
    for (std::size_t i = 0; i < n; ++i) a[i] = b[i] = i;
    std::shuffle(a.begin(), a.end(), gen);
    std::shuffle(b.begin(), b.end(), gen);
    iqs::Permutation p(a, trial % 2 ? "direct" : "inverse");
    print_vec("map", p.map);
    print_vec("imap", p.imap);
    printf("str%s |%s\n", p.GetMapStr().c_str(), p.GetImapStr().c_str());
    std::size_t v = gen() % (std::size_t(1) << n);
    printf("d2p %zu %zu p2d %zu s %s %s find %zu\n", v, p.data2program_(v), p.program2data_(v), p.data2program(v).c_str(),
           p.program2data(v).c_str(), p.Find(n / 2));
    for (std::size_t M = 0; M <= n; M += 1 + n / 3) {
      std::vector<std::size_t> i1, i2;
      p.ObtainIntemediateInverseMaps(b, M, i1, i2);
      printf("M %zu\n", M);
      print_vec(" i1", i1);
      print_vec(" i2", i2);
    }
    p.ExchangeTwoElements(0, n - 1);
    print_vec("xchg", p.map);
    printf("bin %s %zu\n", p.dec2bin(v, n).c_str(), p.bin2dec(p.dec2bin(v, n)));
  }
  iqs::Permutation id(5);
  print_vec("id", id.map);
  printf("op[] %u %d size %zu\n", id[2u], id[3], id.size());

  // --- TinyMatrix ----------------------------------------------------------------------------
  TM2x2<ComplexDP> m;
  m(0, 0) = {1, 2};
  m(0, 1) = {3, 4};
  m[1][0] = {5, 6};
  m[1][1] = {7, 8};
  TM2x2<ComplexDP> c(m);
  printf("tm %d %d rows %u cols %u size %u str %s\n", (int)(c == m), (int)(c != m), m.numRows(), m.numCols(), m.size(), m.tostr().c_str());
  TM4x4<ComplexDP> big;
  for (unsigned i = 0; i < 4; ++i)
    for (unsigned j = 0; j < 4; ++j) big(i, j) = ComplexDP(i, j);
  auto sub = big.getSubMatrix<2, 2>(1, 0, 2, 3);
  printf("sub %g %g %g %g\n", sub(0, 0).real(), sub(0, 1).imag(), sub(1, 0).real(), sub(1, 1).imag());

  // --- bit helpers ---------------------------------------------------------------------------
  printf("bits %u %u %u %d %d %ld\n", iqs::ilog2(1024), iqs::floor_power_of_two(1000), iqs::highestBit(37), (int)iqs::isPowerOf2(64),
         (int)iqs::isPowerOf2(65), iqs::popcnt((uint64_t)0xF0F0F0F0F0F0F0F0ull));
  printf("str %s %s\n", iqs::toString(42).c_str(), iqs::toString(2.5).c_str());

  // --- GateCounter ---------------------------------------------------------------------------
  iqs::GateCounter gc(4);
  gc.OneQubitIncrement(0);
  gc.TwoQubitIncrement(0, 1);
  gc.TwoQubitIncrement(2, 3);
  gc.OneQubitIncrement(3);
  gc.TwoQubitIncrement(1, 2);
  printf("gc %d %d %d depth %d\n", gc.GetTotalGateCount(), gc.GetOneQubitGateCount(), gc.GetTwoQubitGateCount(), gc.GetParallelDepth());

  // --- RandomNumberGenerator -----------------------------------------------------------------
  iqs::RandomNumberGenerator<double> rng;
  rng.SetSeedStreamPtrs(777);
  double u[6];
  rng.UniformRandomNumbers(u, 3, -1., 1., "pool");
  rng.UniformRandomNumbers(u + 3, 3, 0., 1., "local");
  printf("rng %.17g %.17g %.17g %.17g %.17g %.17g\n", u[0], u[1], u[2], u[3], u[4], u[5]);
  rng.SkipAhead(1000, "state");
  double g2[2];
  rng.GaussianRandomNumbers(g2, 2, "state");
  int ints[5];
  rng.RandomIntegersInRange(ints, 5, 3, 11, "local");
  printf("rng2 %.17g %.17g ints %d %d %d %d %d counters %zu %zu %zu\n", g2[0], g2[1], ints[0], ints[1], ints[2], ints[3], ints[4],
         rng.GetNumGeneratedOrSkippedLocalNumbers(), rng.GetNumGeneratedOrSkippedStateNumbers(), rng.GetNumGeneratedOrSkippedPoolNumbers());
  iqs::RandomNumberGenerator<double> copy(&rng);
  double x1, x2;
  rng.UniformRandomNumbers(&x1, 1, 0., 1., "state");
  copy.UniformRandomNumbers(&x2, 1, 0., 1., "state");
  printf("rngcopy %d\n", (int)(x1 == x2));
  std::vector<int> arr = {0, 1, 2, 3, 4, 5, 6, 7};
  iqs::ShuffleFisherYates<int, double>(arr, &rng, "local");
  printf("shuffle %d %d %d %d %d %d %d %d\n", arr[0], arr[1], arr[2], arr[3], arr[4], arr[5], arr[6], arr[7]);

  // --- Environment statics without Init (single process) -------------------------------------
  printf("env %d %d %d %d %d\n", iqs::mpi::Environment::GetStateRank(), iqs::mpi::Environment::GetStateSize(), iqs::mpi::Environment::GetPoolRank(),
         iqs::mpi::Environment::GetNumStates(), iqs::mpi::Environment::GetStateId());
  printf("spec %d %d\n", (int)iqs::ConvertSpec2to1(iqs::GateSpec2Q::CRotationY), (int)iqs::ConvertSpec2to1(iqs::GateSpec2Q::CPhase));
  return 0;
}
