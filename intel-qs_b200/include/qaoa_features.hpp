// qaoa_features.hpp -- QAOA helpers (interface of reference include/qaoa_features.hpp:22-104).
// A second QubitRegister (`diag`) is used as a large real vector holding a classical cost
// function; on this engine the loops over psi[i] / diag[i] are device kernels (csrc/kernels_qaoa.cu).
#ifndef QAOA_EXTRA_FEATURES_HPP
#define QAOA_EXTRA_FEATURES_HPP

#include <vector>

#include "qureg.hpp"

namespace iqs {
namespace qaoa {

// diag[z] = number of edges cut by the bipartition z (z in program-qubit order); returns the max cut
template <typename Type>
int InitializeVectorAsMaxCutCostFunction(QubitRegister<Type> &diag, std::vector<int> &adjacency);

// weighted variant; returns the largest cut weight
template <typename Type>
typename QubitRegister<Type>::BaseType InitializeVectorAsWeightedMaxCutCostFunction(
    QubitRegister<Type> &diag, std::vector<typename QubitRegister<Type>::BaseType> &adjacency);

// |psi> <- exp(-i gamma C) |psi>, C = diag
template <typename Type>
void ImplementQaoaLayerBasedOnCostFunction(QubitRegister<Type> &psi, QubitRegister<Type> &diag,
                                           typename QubitRegister<Type>::BaseType gamma);

template <typename Type>
typename QubitRegister<Type>::BaseType GetExpectationValueFromCostFunction(const QubitRegister<Type> &psi, const QubitRegister<Type> &diag);

template <typename Type>
typename QubitRegister<Type>::BaseType GetExpectationValueSquaredFromCostFunction(const QubitRegister<Type> &psi, const QubitRegister<Type> &diag);

template <typename Type>
std::vector<typename QubitRegister<Type>::BaseType> GetHistogramFromCostFunction(const QubitRegister<Type> &psi, const QubitRegister<Type> &diag,
                                                                                 int max_value);
template <typename Type>
std::vector<typename QubitRegister<Type>::BaseType> GetHistogramFromCostFunctionWithWeightsRounded(const QubitRegister<Type> &psi,
                                                                                                   const QubitRegister<Type> &diag, double max_value);
template <typename Type>
std::vector<typename QubitRegister<Type>::BaseType> GetHistogramFromCostFunctionWithWeightsBinned(const QubitRegister<Type> &psi,
                                                                                                  const QubitRegister<Type> &diag, double max_value,
                                                                                                  double bin_width);
}  // namespace qaoa
}  // namespace iqs
#endif
