// qaoa_check.cpp -- a QAOA MaxCut scenario through the public API (QubitRegister + iqs::qaoa::*),
// printing every observable.  TEST INFRASTRUCTURE: built against the reference (its qaoa_features.cpp
// + libiqs_ref) and against the B200 library; the outputs must agree (tests/test_examples_dropin_gpu.py).
#include <cstdio>
#include <vector>

#include "qaoa_features.hpp"
#include "qureg.hpp"

int main(int argc, char **argv) {
  iqs::mpi::Environment env(argc, argv, false);
  if (!env.IsUsefulRank()) return 0;
  const int n = 10;
  // ring + chords
  std::vector<int> adj(n * n, 0);
  auto edge = [&](int a, int b) { adj[a * n + b] = adj[b * n + a] = 1; };
  for (int v = 0; v < n; ++v) edge(v, (v + 1) % n);
  edge(0, 5); edge(2, 7); edge(3, 8); edge(1, 6);
  iqs::QubitRegister<ComplexDP> diag(n, "base", 0), psi(n, "++++", 0);
  int max_cut = iqs::qaoa::InitializeVectorAsMaxCutCostFunction(diag, adj);
  printf("max_cut %d\n", max_cut);
  printf("diag %g %g %g %g\n", diag[0].real(), diag[1].real(), diag[341].real(), diag[682].real());
  double gammas[3] = {0.4, 0.7, 1.1}, betas[3] = {0.8, 0.5, 0.3};
  for (int p = 0; p < 3; ++p) {
    iqs::qaoa::ImplementQaoaLayerBasedOnCostFunction(psi, diag, gammas[p]);
    for (int q = 0; q < n; ++q) psi.ApplyRotationX(q, betas[p]);
    double e = iqs::qaoa::GetExpectationValueFromCostFunction(psi, diag);
    double e2 = iqs::qaoa::GetExpectationValueSquaredFromCostFunction(psi, diag);
    printf("layer %d  <C> %.12f  <C^2> %.12f  norm %.12f\n", p, e, e2, psi.ComputeNorm());
  }
  std::vector<double> h = iqs::qaoa::GetHistogramFromCostFunction(psi, diag, max_cut);
  for (std::size_t k = 0; k < h.size(); ++k) printf("hist %zu %.12f\n", k, h[k]);
  printf("amp %.12f %.12f  %.12f %.12f\n", psi[3].real(), psi[3].imag(), psi[1000].real(), psi[1000].imag());

  // weighted graph, permuted qubit order
  std::vector<double> w(n * n, 0.);
  for (int a = 0; a < n; ++a)
    for (int b = a + 1; b < n; ++b)
      if ((a * 7 + b * 3) % 4 == 0) w[a * n + b] = w[b * n + a] = 0.25 * ((a + 2 * b) % 5 + 1);
  iqs::QubitRegister<ComplexDP> wdiag(n, "base", 0), phi(n, "++++", 0);
  std::vector<std::size_t> map = {3, 0, 7, 1, 9, 2, 5, 4, 8, 6};
  wdiag.PermuteQubits(map, "direct");
  phi.PermuteQubits(map, "direct");
  double wmax = iqs::qaoa::InitializeVectorAsWeightedMaxCutCostFunction(wdiag, w);
  printf("weighted max %.12f  diag %.12f %.12f %.12f\n", wmax, wdiag[5].real(), wdiag[77].real(), wdiag[1023].real());
  iqs::qaoa::ImplementQaoaLayerBasedOnCostFunction(phi, wdiag, 0.37);
  for (int q = 0; q < n; ++q) phi.ApplyRotationX(q, 0.61);
  printf("weighted <C> %.12f <C^2> %.12f\n", iqs::qaoa::GetExpectationValueFromCostFunction(phi, wdiag),
         iqs::qaoa::GetExpectationValueSquaredFromCostFunction(phi, wdiag));
  std::vector<double> hr = iqs::qaoa::GetHistogramFromCostFunctionWithWeightsRounded(phi, wdiag, wmax);
  for (std::size_t k = 0; k < hr.size(); ++k) printf("hist_rounded %zu %.12f\n", k, hr[k]);
  std::vector<double> hb = iqs::qaoa::GetHistogramFromCostFunctionWithWeightsBinned(phi, wdiag, wmax, 0.5);
  for (std::size_t k = 0; k < hb.size(); ++k) printf("hist_binned %zu %.12f\n", k, hb[k]);
  return 0;
}
