// chi_matrix.hpp -- chi-matrix container for quantum channels (interface subset of reference
// include/chi_matrix.hpp).  Channels are clients of the gate path and outside the B200 scope
// (SURVEY.md section 2, row 18): the container exists so that code naming CM4x4 / CM16x16
// compiles; the eigen-decomposition needs Eigen, which is not available offline, and throws.
#ifndef IQS_CHI_MATRIX_HPP
#define IQS_CHI_MATRIX_HPP
#include <stdexcept>
#include <vector>

#include "tinymatrix.hpp"
#include "utils.hpp"

namespace iqs {
template <class ValueType, unsigned M, unsigned align = alignof(ValueType)>
class ChiMatrix : public TinyMatrix<ValueType, M, M, align> {
 public:
  using value_type = ValueType;
  using base_type = typename extract_value_type<ValueType>::value_type;
  using size_type = unsigned;
  ChiMatrix() : TinyMatrix<ValueType, M, M, align>() {}
  void SolveEigenSystem() { throw std::runtime_error("ChiMatrix::SolveEigenSystem: quantum channels are outside the scope of the B200 engine"); }
  void EigensystemOfIdealHadamardChannel() { SolveEigenSystem(); }
  value_type GetEigenValue(size_type k) const { return evalues_.at(k); }
  std::vector<value_type> GetEigenValues() const { return evalues_; }
  base_type GetEigenProbability(size_type k) const { return eprobs_.at(k); }
  std::vector<base_type> GetEigenProbabilities() const { return eprobs_; }
  base_type GetEigenCumulativeProbability(size_type k) const { return ecumprobs_.at(k); }
  std::vector<base_type> GetEigenCumulativeProbabilities() const { return ecumprobs_; }
  std::vector<value_type> GetEigenVector(size_type k) const { return evectors_.at(k); }
  std::vector<std::vector<value_type>> GetEigenVectors() const { return evectors_; }
  void Print() { this->print("chi"); }

 private:
  std::vector<value_type> evalues_;
  std::vector<base_type> eprobs_, ecumprobs_;
  std::vector<std::vector<value_type>> evectors_;
};
}  // namespace iqs
#endif
