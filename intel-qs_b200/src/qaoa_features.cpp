// qaoa_features.cpp -- host side of the QAOA helpers: argument checks as in the reference
// (src/qaoa_features.cpp), the loops run on the device through the C ABI (iqsb_qaoa_*).
#include "../include/qaoa_features.hpp"

#include <cassert>
#include <cmath>

#include "qureg_impl.hpp"

namespace iqs {
namespace qaoa {

using detail::Check;

namespace {
template <typename Type>
void SameShape(const QubitRegister<Type> &psi, const QubitRegister<Type> &diag) {
  assert(psi.LocalSize() == diag.LocalSize());
  assert(psi.GlobalSize() == diag.GlobalSize());
  assert(psi.qubit_permutation->map == diag.qubit_permutation->map);
  psi.PrepareDevice();
  diag.PrepareDevice();
}

template <typename Type>
double FillCuts(QubitRegister<Type> &diag, const std::vector<double> &adj, bool weighted) {
  unsigned n = (unsigned)diag.NumQubits();
  assert(adj.size() == std::size_t(n) * n);
  for (unsigned v = 0; v < n; ++v) assert(adj[v * n + v] == 0);
  std::vector<uint8_t> pos(n);
  for (unsigned q = 0; q < n; ++q) pos[q] = (uint8_t)diag.qubit_permutation->map[q];
  diag.PrepareDevice();
  std::size_t glb_start = UL(iqs::mpi::Environment::GetStateRank()) * diag.LocalSize();
  double mx = 0;
  Check(iqsb_qaoa_maxcut(diag.DeviceState(), n, adj.data(), weighted ? 1 : 0, pos.data(), glb_start, &mx), "MaxCut cost function");
  iqs::mpi::AllreduceDouble(&mx, 1, iqs::mpi::MAX);
  return mx;
}

template <typename Type>
std::vector<typename QubitRegister<Type>::BaseType> Histogram(const QubitRegister<Type> &psi, const QubitRegister<Type> &diag, int nbins, double width,
                                                              double eps) {
  SameShape(psi, diag);
  std::vector<double> h(nbins, 0.);
  Check(iqsb_qaoa_histogram(psi.DeviceState(), diag.DeviceState(), nbins, width, eps, h.data()), "cost-function histogram");
  for (int first = 0; first < nbins; first += 32) iqs::mpi::AllreduceDouble(h.data() + first, std::min(32, nbins - first), iqs::mpi::SUM);
  return std::vector<typename QubitRegister<Type>::BaseType>(h.begin(), h.end());
}
}  // namespace

template <typename Type>
int InitializeVectorAsMaxCutCostFunction(QubitRegister<Type> &diag, std::vector<int> &adjacency) {
  long total = 0;
  for (int e : adjacency) total += e;
  assert(total % 2 == 0);
  (void)total;
  std::vector<double> adj(adjacency.begin(), adjacency.end());
  return (int)FillCuts(diag, adj, false);
}

template <typename Type>
typename QubitRegister<Type>::BaseType InitializeVectorAsWeightedMaxCutCostFunction(
    QubitRegister<Type> &diag, std::vector<typename QubitRegister<Type>::BaseType> &adjacency) {
  unsigned n = (unsigned)diag.NumQubits();
  for (unsigned v1 = 0; v1 < n; ++v1)
    for (unsigned v2 = v1 + 1; v2 < n; ++v2) assert(adjacency[v1 * n + v2] == adjacency[v2 * n + v1]);
  (void)n;
  std::vector<double> adj(adjacency.begin(), adjacency.end());
  return (typename QubitRegister<Type>::BaseType)FillCuts(diag, adj, true);
}

template <typename Type>
void ImplementQaoaLayerBasedOnCostFunction(QubitRegister<Type> &psi, QubitRegister<Type> &diag, typename QubitRegister<Type>::BaseType gamma) {
  SameShape(psi, diag);
  Check(iqsb_qaoa_layer(psi.DeviceState(), diag.DeviceState(), (double)gamma), "QAOA layer");
}

template <typename Type>
typename QubitRegister<Type>::BaseType GetExpectationValueFromCostFunction(const QubitRegister<Type> &psi, const QubitRegister<Type> &diag) {
  SameShape(psi, diag);
  double v[2];
  Check(iqsb_qaoa_expect(psi.DeviceState(), diag.DeviceState(), v), "cost-function expectation");
  iqs::mpi::AllreduceDouble(v, 2, iqs::mpi::SUM);
  return (typename QubitRegister<Type>::BaseType)v[0];
}

template <typename Type>
typename QubitRegister<Type>::BaseType GetExpectationValueSquaredFromCostFunction(const QubitRegister<Type> &psi, const QubitRegister<Type> &diag) {
  SameShape(psi, diag);
  double v[2];
  Check(iqsb_qaoa_expect(psi.DeviceState(), diag.DeviceState(), v), "cost-function expectation");
  iqs::mpi::AllreduceDouble(v, 2, iqs::mpi::SUM);
  return (typename QubitRegister<Type>::BaseType)v[1];
}

template <typename Type>
std::vector<typename QubitRegister<Type>::BaseType> GetHistogramFromCostFunction(const QubitRegister<Type> &psi, const QubitRegister<Type> &diag,
                                                                                 int max_value) {
  assert(max_value > 0);
  return Histogram(psi, diag, max_value + 1, 1.0, 0.0);  // bin = (int) cut  (:371-374)
}

template <typename Type>
std::vector<typename QubitRegister<Type>::BaseType> GetHistogramFromCostFunctionWithWeightsRounded(const QubitRegister<Type> &psi,
                                                                                                   const QubitRegister<Type> &diag, double max_value) {
  assert(max_value > 0);
  return Histogram(psi, diag, (int)(std::floor(max_value)) + 1, 1.0, 1e-7);  // bin = floor(cut + 1e-7)  (:430-433)
}

template <typename Type>
std::vector<typename QubitRegister<Type>::BaseType> GetHistogramFromCostFunctionWithWeightsBinned(const QubitRegister<Type> &psi,
                                                                                                  const QubitRegister<Type> &diag, double max_value,
                                                                                                  double bin_width) {
  assert(max_value > 0);
  return Histogram(psi, diag, (int)(std::ceil(max_value / bin_width)) + 1, bin_width, 1e-7);  // bin = floor(cut / width + 1e-7)  (:493-496)
}

#define IQS_QAOA_INSTANTIATE(Type, Base)                                                                                                       \
  template int InitializeVectorAsMaxCutCostFunction<Type>(QubitRegister<Type> &, std::vector<int> &);                                          \
  template Base InitializeVectorAsWeightedMaxCutCostFunction<Type>(QubitRegister<Type> &, std::vector<Base> &);                                \
  template void ImplementQaoaLayerBasedOnCostFunction<Type>(QubitRegister<Type> &, QubitRegister<Type> &, Base);                               \
  template Base GetExpectationValueFromCostFunction<Type>(const QubitRegister<Type> &, const QubitRegister<Type> &);                           \
  template Base GetExpectationValueSquaredFromCostFunction<Type>(const QubitRegister<Type> &, const QubitRegister<Type> &);                    \
  template std::vector<Base> GetHistogramFromCostFunction<Type>(const QubitRegister<Type> &, const QubitRegister<Type> &, int);                \
  template std::vector<Base> GetHistogramFromCostFunctionWithWeightsRounded<Type>(const QubitRegister<Type> &, const QubitRegister<Type> &, double); \
  template std::vector<Base> GetHistogramFromCostFunctionWithWeightsBinned<Type>(const QubitRegister<Type> &, const QubitRegister<Type> &, double, double);
IQS_QAOA_INSTANTIATE(ComplexDP, double)
IQS_QAOA_INSTANTIATE(ComplexSP, float)

}  // namespace qaoa
}  // namespace iqs
