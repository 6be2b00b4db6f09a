// noise_check.cpp -- noisy simulation through the public API: NoisyQureg, ApplyNoiseGate and
// ApplyChannel with the closed-form eigensystem of the ideal Hadamard channel.  TEST
// INFRASTRUCTURE: built against the reference (libiqs_ref) and against the B200 library; the
// outputs must agree (tests/test_examples_dropin_gpu.py).  The part under IQS_WITH_NOISE needs the
// chi-matrix eigen-solver, which the reference only has when configured with Eigen: it is built
// for the B200 library only and checks physical invariants instead of reference output.
#include <cstdio>
#include <vector>

#include "qureg.hpp"

template <class R>
static void dump(const char *tag, R &psi, std::initializer_list<std::size_t> idx) {
  printf("%s norm %.13f |", tag, psi.ComputeNorm());
  for (auto i : idx) {
    ComplexDP a = psi.GetGlobalAmplitude(i);
    printf(" [%zu] %.13f %.13f", i, a.real(), a.imag());
  }
  printf("\n");
}

int main(int argc, char **argv) {
  iqs::mpi::Environment env(argc, argv, false);
  if (!env.IsUsefulRank()) return 0;
  const unsigned n = 9;
  TM2x2<ComplexDP> G;
  G(0, 0) = {0.592056606032915, 0.459533060553574};
  G(0, 1) = {-0.314948020757856, -0.582328159830658};
  G(1, 0) = {0.658235557641767, 0.070882241549507};
  G(1, 1) = {0.649564427121402, 0.373855203932477};

  // 1. NoisyQureg: every overridden gate, counters, durations, both noise-gate constructions
  {
    iqs::NoisyQureg<ComplexDP> psi(n, 4242, 300., 120.);
    psi.Initialize("base", 5);
    psi.SetGateDurations(1.5, 4.);
    for (unsigned q = 0; q < n; ++q) psi.ApplyHadamard(q);
    dump("noisy H", psi, {0, 5, 100, 511});
    for (unsigned q = 0; q + 1 < n; ++q) psi.ApplyCPauliX(q, q + 1);
    psi.ApplyRotationX(2, 0.3);
    psi.ApplyRotationY(7, 1.3);
    psi.ApplyRotationZ(0, 2.1);
    psi.Apply1QubitGate(4, G);
    psi.ApplyControlled1QubitGate(8, 1, G);
    dump("noisy circuit", psi, {1, 77, 300, 510});
    psi.ApplyNoiseGatesOnAllQubits();
    dump("noisy final", psi, {1, 77, 300, 510});
    printf("counts total %u one %u two %u  (3,3) %u (3,4) %u (1,8) %u\n", psi.GetTotalExperimentalGateCount(), psi.GetOneQubitExperimentalGateCount(),
           psi.GetTwoQubitExperimentalGateCount(), psi.GetExperimentalGateCount(3, 3), psi.GetExperimentalGateCount(3, 4), psi.GetExperimentalGateCount(1, 8));
    std::vector<unsigned> row = psi.GetExperimentalGateCount(4);
    printf("row 4:");
    for (unsigned c : row) printf(" %u", c);
    printf("\n");
    psi.SetDecoherenceTime(50., 30.);
    psi.ApplyHadamard(3);
    psi.ApplyHadamard(5);
    psi.NoiseGate_OLD(1);
    psi.NoiseGate(6);
    psi.AddNoiseOneQubitGate(2);
    psi.AddNoiseTwoQubitGate(0, 8);
    dump("noisy old+new", psi, {0, 64, 255, 256});
    psi.Initialize("base", 0);
    printf("after Initialize: total %u\n", psi.GetTotalExperimentalGateCount());
    psi.ApplyHadamard(0);  // no time has passed: no noise gate
    dump("fresh H", psi, {0, 1});
  }

  // 2. ApplyNoiseGate of the base class (RNG stream "state"), as the noisy tutorial uses it
  {
    iqs::RandomNumberGenerator<double> rng;
    rng.SetSeedStreamPtrs(31415);
    iqs::QubitRegister<ComplexDP> psi(n, "base", 0);
    psi.SetRngPtr(&rng);
    psi.SetNoiseTimescales(40., 20.);
    printf("T1 %g T2 %g Tphi %.12f\n", psi.GetT1(), psi.GetT2(), psi.GetTphi());
    for (unsigned q = 0; q < n; ++q) {
      psi.ApplyRotationY(q, 0.4 + 0.1 * q);
      for (unsigned p = 0; p < n; ++p) psi.ApplyNoiseGate(p, 1.5);
    }
    dump("noise gates", psi, {0, 3, 200, 511});
    printf("P(q0) %.13f P(q8) %.13f\n", psi.GetProbability(0), psi.GetProbability(8));
  }

  // 3. channel with a known eigensystem: the ideal Hadamard as chi matrix
  {
    CM4x4<ComplexDP> chi;
    chi(1, 1) = chi(1, 3) = chi(3, 1) = chi(3, 3) = ComplexDP(0.5, 0);
    chi.EigensystemOfIdealHadamardChannel();
    printf("chi(H): E %.12f %.12f %.12f %.12f  cum %.12f %.12f\n", chi.GetEigenValue(0).real(), chi.GetEigenValue(1).real(), chi.GetEigenValue(2).real(),
           chi.GetEigenValue(3).real(), chi.GetEigenCumulativeProbability(0), chi.GetEigenCumulativeProbability(3));
    std::vector<ComplexDP> e0 = chi.GetEigenVector(0);
    printf("chi(H): |E0> %.12f %.12f %.12f %.12f  p0 %.12f\n", e0[0].real(), e0[1].real(), e0[2].real(), e0[3].real(), chi.GetEigenProbability(0));
    iqs::RandomNumberGenerator<double> rng;
    rng.SetSeedStreamPtrs(7777);
    iqs::QubitRegister<ComplexDP> psi(n, "base", 4), ideal(n, "base", 4);
    psi.SetRngPtr(&rng);
    for (unsigned q = 0; q < n; ++q) {
      psi.ApplyChannel(q, chi);
      ideal.ApplyHadamard(q);
    }
    psi.ApplyChannel(2, chi);
    ideal.ApplyHadamard(2);
    dump("channel H", psi, {0, 4, 100, 511});
    printf("overlap with ideal %.13f  sign %g\n", std::norm(ideal.ComputeOverlap(psi)), psi.GetOverallSignOfChannels());
    CM4x4<ComplexDP> copy(chi);
    iqs::ChiMatrix<ComplexDP, 4> unaligned(chi);
    printf("copies: %d %d %.12f\n", copy == chi, unaligned == chi, unaligned.GetEigenVector(1)[3].real());
  }

#ifdef IQS_WITH_NOISE
  printf("==== eigen-solver part (not built for the reference without Eigen)\n");
  {
    // depolarising channel rho' = (1-p) rho + p/3 (X rho X + Y rho Y + Z rho Z): the overlap with the
    // initial state decays, the norm of every trajectory stays 1 (the eigen-operators are scaled Paulis)
    double p = 0.01;
    CM4x4<ComplexDP> chi;
    for (int i = 0; i < 4; ++i) chi(i, i) = ComplexDP(i == 0 ? 1 - p : p / 3, 0);
    chi.SolveEigenSystem();
    iqs::RandomNumberGenerator<double> rng;
    rng.SetSeedStreamPtrs(7777);
    iqs::QubitRegister<ComplexDP> psi0(4, "base", 1);
    psi0.ApplyHadamard(1);
    const int steps = 20, ensemble = 100;
    std::vector<double> ov(steps, 0.);
    double worst_norm = 0;
    for (int s = 0; s < ensemble; ++s) {
      iqs::QubitRegister<ComplexDP> psi(psi0);
      psi.SetRngPtr(&rng);
      ov[0] += std::norm(psi0.ComputeOverlap(psi));
      for (int t = 1; t < steps; ++t) {
        for (unsigned q = 0; q < 4; ++q) psi.ApplyChannel(q, chi);
        ov[t] += std::norm(psi0.ComputeOverlap(psi));
      }
      worst_norm = std::max(worst_norm, std::abs(psi.ComputeNorm() - 1.));
    }
    // exact: each qubit keeps its state with probability (1 - 4p/3 ...) -- for this product state the
    // overlap after t steps is prod over qubits of (1 - 2p/3 (1 - ...)); checked loosely against 1 - 4*(2p/3)*t
    printf("depolarising: ov[0] %.6f ov[10] %.3f ov[19] %.3f  max |norm-1| %.2e\n", ov[0] / ensemble, ov[10] / ensemble, ov[19] / ensemble, worst_norm);
    bool ok = std::abs(ov[0] / ensemble - 1.) < 1e-12 && ov[19] / ensemble < 0.95 && ov[19] / ensemble > 0.3 && worst_norm < 1e-12;
    printf("depolarising %s\n", ok ? "OK" : "FAILED");
  }
  {
    // two-qubit channel of the ideal CZ gate: CZ = 1/2 (id.id + id.Z + Z.id - Z.Z), chi = |v><v| with
    // v = 1/2 (1,0,0,1, 0,..., 1,0,0,-1) on {id.id, id.Z, Z.id, Z.Z} = indices 0, 3, 12, 15
    CM16x16<ComplexDP> chi;
    const int idx[4] = {0, 3, 12, 15};
    const double v[4] = {0.5, 0.5, 0.5, -0.5};
    for (int a = 0; a < 4; ++a)
      for (int b = 0; b < 4; ++b) chi(idx[a], idx[b]) = ComplexDP(v[a] * v[b], 0);
    chi.SolveEigenSystem();
    iqs::RandomNumberGenerator<double> rng;
    rng.SetSeedStreamPtrs(99);
    iqs::QubitRegister<ComplexDP> psi(6, "++++", 0), ideal(6, "++++", 0);
    psi.SetRngPtr(&rng);
    psi.ApplyRotationY(1, 0.7);
    ideal.ApplyRotationY(1, 0.7);
    psi.ApplyChannel(1, 4, chi);
    ideal.ApplyCPauliZ(1, 4);
    psi.ApplyChannel(5, 0, chi);
    ideal.ApplyCPauliZ(5, 0);
    double ov = std::norm(ideal.ComputeOverlap(psi));
    printf("CZ channel: |<ideal|psi>|^2 %.13f  %s\n", ov, std::abs(ov - 1.) < 1e-12 ? "OK" : "FAILED");
  }
#endif
  return 0;
}
