// pool_check.cpp -- pool of states with one state per GPU (iqs::mpi::Environment::UpdateStateComm),
// restating the reference's unit_test/include/noisy_simulation_test.hpp:52-158 and
// multiple_states_test.hpp for the split this engine supports.  TEST INFRASTRUCTURE: launched by
// tests/test_multigpu.py with tools/iqsrun -n 2; prints "OK <name>" lines, exits 1 on a failure.
#include <cmath>
#include <cstdio>

#include "qureg.hpp"

static int failures = 0;
#define EXPECT(cond, name)                                                                   \
  do {                                                                                       \
    if (!(cond)) { printf("FAILED %s: %s (pool rank %d)\n", name, #cond, iqs::mpi::Environment::GetPoolRank()); ++failures; } \
  } while (0)

int main(int argc, char **argv) {
  using Env = iqs::mpi::Environment;
  Env env(argc, argv, false);
  const int world = Env::GetPoolSize(), me = Env::GetPoolRank();
  const int n = 6;
  const double T1 = 6., T2 = 4.;
  EXPECT(Env::GetNumStates() == 1 && Env::GetStateSize() == world && Env::GetStateRank() == me, "one state over all ranks");

  // one state at a time (noisy_simulation_test.hpp:52-72): the register is sharded over all GPUs
  {
    iqs::QubitRegister<ComplexDP> psi(n, "base", 1 + 8 + 16 + 32);
    psi.ApplyHadamard(0);
    psi.ApplyHadamard(1);
    iqs::QubitRegister<ComplexDP> noisy(psi);
    EXPECT(std::abs(noisy.ComputeOverlap(psi).real() - 1.) < 1e-14, "copy");
    noisy.SetNoiseTimescales(T1, T2);
    iqs::RandomNumberGenerator<double> rng;
    rng.SetSeedStreamPtrs(7777);
    noisy.SetRngPtr(&rng);
    for (int q = 0; q < n; ++q) noisy.ApplyNoiseGate(q, 5.);
    EXPECT(noisy.ComputeOverlap(psi).real() < 1. - 1e-15, "noise changes the state");
    EXPECT(std::abs(noisy.ComputeNorm() - 1.) < 1e-13, "noise gates are unitary");
  }
  if (me == 0) printf("OK one_state_at_a_time\n");

  // one state per rank (:123-158)
  env.UpdateStateComm(world);
  EXPECT(Env::GetNumStates() == world && Env::GetStateRank() == 0 && Env::GetStateSize() == 1, "split");
  EXPECT(Env::GetPoolRank() == me && Env::GetPoolSize() == world && Env::GetStateId() == me && Env::IsUsefulRank(), "ids");
  {
    iqs::QubitRegister<ComplexDP> psi(n, "base", 0);
    EXPECT(psi.GlobalSize() == psi.LocalSize(), "every state is local to its GPU");
    std::size_t index = me % psi.GlobalSize();
    psi.Initialize("base", index);
    double p = psi.GetProbability(0);
    EXPECT(p == double(index % 2), "probability of the own state");
    double sum = Env::IncoherentSumOverAllStatesOfPool<double>(p);
    EXPECT(sum == double(world / 2), "incoherent sum over the pool");

    // an ensemble of noisy trajectories, different on every GPU (state stream of the RNG)
    iqs::RandomNumberGenerator<double> rng;
    rng.SetSeedStreamPtrs(7777);
    psi.Initialize("base", 0);
    for (int q = 0; q < n; ++q) psi.ApplyHadamard(q);
    iqs::QubitRegister<ComplexDP> noisy(psi);
    noisy.SetRngPtr(&rng);
    noisy.SetNoiseTimescales(T1, T1 / 2);
    for (int q = 0; q < n; ++q) noisy.ApplyNoiseGate(q, T1);
    double ov = std::norm(noisy.ComputeOverlap(psi));
    double ov_sum = Env::IncoherentSumOverAllStatesOfPool<double>(ov);
    double ov_sq_sum = Env::IncoherentSumOverAllStatesOfPool<double>(ov * ov);
    EXPECT(ov < 1. && ov > 0., "trajectory decohered");
    EXPECT(ov_sum < world && ov_sum > 0., "ensemble sum");
    // the trajectories differ between GPUs: the variance over the pool is not zero
    double mean = ov_sum / world, var = ov_sq_sum / world - mean * mean;
    EXPECT(var > 1e-12, "different noise on different GPUs");
    double avg_p = Env::IncoherentSumOverAllStatesOfPool<double>(noisy.GetProbability(0)) / world;
    EXPECT(avg_p > 1e-15 && avg_p < 1., "average probability");
    iqs::mpi::PoolBarrier();
    if (me == 0) printf("OK one_state_per_rank  <ov> = %.6f  var = %.3e\n", mean, var);
  }

  // an unsupported split is refused loudly; back to one state
  if (world > 2) {
    bool threw = false;
    try { env.UpdateStateComm(2); } catch (std::runtime_error const &) { threw = true; }
    EXPECT(threw, "intermediate splits are refused");
  }
  env.UpdateStateComm(1);
  EXPECT(Env::GetNumStates() == 1 && Env::GetStateSize() == world, "back to one state");
  {
    iqs::QubitRegister<ComplexDP> psi(n, "++++", 0);
    psi.ApplyHadamard(n - 1);  // a global qubit again
    EXPECT(std::abs(psi.GetProbability(n - 1)) < 1e-14 && std::abs(psi.ComputeNorm() - 1.) < 1e-14, "sharded register after the pool");
  }
  double total_failures = Env::IncoherentSumOverAllStatesOfPool<double>(double(failures));
  if (me == 0) printf(total_failures == 0 ? "ALL OK\n" : "FAILED\n");
  return total_failures == 0 ? 0 : 1;
}
