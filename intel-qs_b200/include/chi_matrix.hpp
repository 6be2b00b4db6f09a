// chi_matrix.hpp -- chi-matrix of a quantum channel with its eigen-decomposition
// (interface of reference include/chi_matrix.hpp:46-205, behaviour of src/chi_matrix.cpp:58-240).
//
//   rho' = sum_ij chi_ij sigma_i rho sigma_j^dagger,   Pauli basis {id, X, Y, Z} (two qubits:
//   {id.id, id.X, id.Y, id.Z, X.id, ...}),   chi = sum_k E_k |E_k><E_k|.
//
// QubitRegister::ApplyChannel samples k with probability |E_k| / sum|E_k| and applies the operator
// sum_i E_k,i sigma_i, so the eigenvectors are kept "standardised" (chi = sum_k E_k |E_k><E_k|
// holds exactly, chi_matrix.cpp:85-118) and "renormalised" by sqrt(sum_k |E_k|)
// (chi_matrix.cpp:177-207).
//
// The reference delegates the decomposition to Eigen's ComplexEigenSolver (a configure-time
// download that is not part of its source tree); here it is a cyclic complex Jacobi iteration for
// Hermitian matrices, M <= 16, written for this file.  Conventions chosen so that the reference's
// known-answer test holds (unit_test/include/chi_matrix_test.hpp:139-163): eigenvalues ascending,
// each unit eigenvector phased so that its first non-negligible component is real and negative.
// For a degenerate eigenvalue any orthonormal basis of the eigenspace is a valid answer, so the
// sampled sequence of operators may differ from an Eigen build while the channel is the same.
#ifndef IQS_CHI_MATRIX_HPP
#define IQS_CHI_MATRIX_HPP
#include <algorithm>
#include <cassert>
#include <cmath>
#include <complex>
#include <initializer_list>
#include <iostream>
#include <limits>
#include <stdexcept>
#include <vector>

#include "tinymatrix.hpp"
#include "utils.hpp"

namespace iqs {

namespace detail {
// Hermitian eigenproblem A = V diag(w) V^H by cyclic Jacobi rotations; a is M x M row-major and is
// destroyed, v receives the eigenvectors as columns (row-major).
template <class C>
void HermitianJacobi(unsigned M, std::vector<C> &a, std::vector<C> &v) {
  using R = typename C::value_type;
  v.assign((size_t)M * M, C(0));
  for (unsigned i = 0; i < M; ++i) v[(size_t)i * M + i] = C(1);
  for (int sweep = 0; sweep < 64; ++sweep) {
    R off = 0, diag = 0;
    for (unsigned i = 0; i < M; ++i)
      for (unsigned j = 0; j < M; ++j) (i == j ? diag : off) += std::norm(a[(size_t)i * M + j]);
    if (off <= std::numeric_limits<R>::epsilon() * std::numeric_limits<R>::epsilon() * (diag + off) || off == 0) break;
    for (unsigned p = 0; p + 1 < M; ++p)
      for (unsigned q = p + 1; q < M; ++q) {
        const C apq = a[(size_t)p * M + q];
        const R mag = std::abs(apq);
        if (mag == 0) continue;
        const C phase = apq / mag;  // e^{i phi}
        const R tau = (a[(size_t)q * M + q].real() - a[(size_t)p * M + p].real()) / (2 * mag);
        const R t = (tau >= 0 ? R(1) : R(-1)) / (std::abs(tau) + std::sqrt(1 + tau * tau));
        const R c = 1 / std::sqrt(1 + t * t), s = t * c;
        // J = [[c, s], [-s e^{-i phi}, c e^{-i phi}]] on (p, q):  A <- J^H A J,  V <- V J
        const C jpp = c, jpq = s, jqp = -s * std::conj(phase), jqq = c * std::conj(phase);
        for (unsigned k = 0; k < M; ++k) {  // columns p, q of A and V
          C akp = a[(size_t)k * M + p], akq = a[(size_t)k * M + q];
          a[(size_t)k * M + p] = akp * jpp + akq * jqp;
          a[(size_t)k * M + q] = akp * jpq + akq * jqq;
          C vkp = v[(size_t)k * M + p], vkq = v[(size_t)k * M + q];
          v[(size_t)k * M + p] = vkp * jpp + vkq * jqp;
          v[(size_t)k * M + q] = vkp * jpq + vkq * jqq;
        }
        for (unsigned k = 0; k < M; ++k) {  // rows p, q of A
          C apk = a[(size_t)p * M + k], aqk = a[(size_t)q * M + k];
          a[(size_t)p * M + k] = std::conj(jpp) * apk + std::conj(jqp) * aqk;
          a[(size_t)q * M + k] = std::conj(jpq) * apk + std::conj(jqq) * aqk;
        }
        a[(size_t)p * M + q] = C(0);
        a[(size_t)q * M + p] = C(0);
        a[(size_t)p * M + p] = C(a[(size_t)p * M + p].real());
        a[(size_t)q * M + q] = C(a[(size_t)q * M + q].real());
      }
  }
}
}  // namespace detail

template <class ValueType, unsigned M, unsigned align = alignof(ValueType)>
class ChiMatrix : public TinyMatrix<ValueType, M, M, align> {
  using Base = TinyMatrix<ValueType, M, M, align>;

 public:
  using value_type = ValueType;
  typedef typename extract_value_type<ValueType>::value_type BaseType;
  using base_type = BaseType;
  using size_type = unsigned;
  using RowType = ValueType[M];

  ChiMatrix() : Base() {
    for (size_type i = 0; i < M; ++i)
      for (size_type j = 0; j < M; ++j) this->data_[i][j] = ValueType();
  }
  template <class U>
  ChiMatrix(U init[M][M]) : Base(init) {}
  template <class U>
  ChiMatrix(std::initializer_list<std::initializer_list<U>> const &init) : Base(init) {}
  // copy from another element type / alignment, eigensystem included (chi_matrix.hpp:86-99)
  template <class U, unsigned alignrhs>
  ChiMatrix(ChiMatrix<U, M, alignrhs> const &rhs) {
    for (size_type i = 0; i < M; ++i)
      for (size_type j = 0; j < M; ++j) this->data_[i][j] = rhs(i, j);
    evalues_ = rhs.GetEigenValues();
    eprobs_ = rhs.GetEigenProbabilities();
    ecumprobs_ = rhs.GetEigenCumulativeProbabilities();
    evectors_ = rhs.GetEigenVectors();
  }
  ChiMatrix(ChiMatrix const &) = default;
  ChiMatrix &operator=(ChiMatrix const &) = default;

  ValueType *GetPtrToData() { return &(this->data_[0][0]); }

  // chi(H) = 1/2 (|X> + |Z>)(<X| + <Z|): eigenvalue 1 for (0,1,0,1)/sqrt(2), 0 three times.
  // Closed form for CM4x4<ComplexDP> (chi_matrix.cpp:212-238); any other instantiation is the
  // reference's placeholder that only prints.
  void EigensystemOfIdealHadamardChannel() { IdealHadamard(static_cast<ValueType *>(nullptr)); }

  // eigenvalues / eigenvectors of the (Hermitian) chi matrix; must be called after the entries are set
  void SolveEigenSystem() { Solve(static_cast<ValueType *>(nullptr)); }

  void Print(bool with_eigensystem = true) {
    std::cout << "chi_matrix :\n";
    for (size_type i = 0; i < M; ++i) {
      for (size_type j = 0; j < M; ++j) std::cout << this->data_[i][j] << "\t";
      std::cout << "\n";
    }
    if (!with_eigensystem) return;
    std::cout << "eigenvalues :\n";
    for (auto const &e : evalues_) std::cout << e << "\t";
    std::cout << "\neigenprobs :\n";
    for (auto const &p : eprobs_) std::cout << p << "\t";
    for (size_type k = 0; k < evectors_.size(); ++k) {
      std::cout << "\neigenvector " << k << " :\n";
      for (auto const &x : evectors_[k]) std::cout << x << "\t";
    }
    std::cout << "\n";
  }

  // unchecked on purpose, like the reference (called once per channel application)
  value_type GetEigenValue(size_type k) const { return evalues_[k]; }
  std::vector<value_type> GetEigenValues() const { return evalues_; }
  base_type GetEigenProbability(size_type k) const { return eprobs_[k]; }
  std::vector<base_type> GetEigenProbabilities() const { return eprobs_; }
  base_type GetEigenCumulativeProbability(size_type k) const { return ecumprobs_[k]; }
  std::vector<base_type> GetEigenCumulativeProbabilities() const { return ecumprobs_; }
  std::vector<value_type> GetEigenVector(size_type k) const { return evectors_[k]; }
  std::vector<std::vector<value_type>> GetEigenVectors() const { return evectors_; }

 protected:
  std::vector<value_type> evalues_;                 // complex type, real values (reference layout)
  std::vector<std::vector<value_type>> evectors_;   // evectors_[k][i] = component i of |E_k>
  std::vector<base_type> eprobs_, ecumprobs_;

  // p(k) = |E_k| / sum |E_k|; eigenvectors scaled by sqrt(sum |E_k|)   (chi_matrix.cpp:177-207)
  void NormalizeEigenProbAndRenormalizeEigenVect() {
    base_type total = 0;
    eprobs_.clear();
    ecumprobs_.clear();
    for (auto const &e : evalues_) {
      assert(std::abs(std::imag(e)) < 1.0e-14 && "Eigenvalues of chi matrix must be real.");
      base_type w = std::abs(std::real(e));
      eprobs_.push_back(w);
      total += w;
      ecumprobs_.push_back(total);
    }
    if (total == 0 || total == 1) return;
    for (size_type k = 0; k < eprobs_.size(); ++k) {
      eprobs_[k] /= total;
      ecumprobs_[k] /= total;
    }
    const base_type scale = std::sqrt(total);
    for (auto &vec : evectors_)
      for (auto &x : vec) x *= scale;
  }

 private:
  template <class R>
  void Solve(std::complex<R> *) {
    using C = std::complex<R>;
    std::vector<C> a((size_t)M * M), v;
    R scale = 0;
    for (size_type i = 0; i < M; ++i)
      for (size_type j = 0; j < M; ++j) scale = std::max(scale, (R)std::abs(this->data_[i][j]));
    for (size_type i = 0; i < M; ++i)
      for (size_type j = 0; j < M; ++j) {
        if (std::abs(this->data_[i][j] - std::conj(this->data_[j][i])) > 1e-6 * scale)
          throw std::invalid_argument("ChiMatrix::SolveEigenSystem: the chi matrix of a channel is Hermitian; this one is not");
        a[(size_t)i * M + j] = (this->data_[i][j] + std::conj(this->data_[j][i])) / R(2);
      }
    detail::HermitianJacobi<C>(M, a, v);
    std::vector<size_type> order(M);
    for (size_type k = 0; k < M; ++k) order[k] = k;
    std::stable_sort(order.begin(), order.end(), [&](size_type x, size_type y) { return a[(size_t)x * M + x].real() < a[(size_t)y * M + y].real(); });
    evalues_.assign(M, C(0));
    evectors_.assign(M, std::vector<C>(M, C(0)));
    for (size_type k = 0; k < M; ++k) {
      const size_type col = order[k];
      evalues_[k] = C(a[(size_t)col * M + col].real());
      std::vector<C> &vec = evectors_[k];
      R nrm = 0;
      for (size_type i = 0; i < M; ++i) {
        vec[i] = v[(size_t)i * M + col];
        nrm += std::norm(vec[i]);
      }
      nrm = std::sqrt(nrm);
      C ph(-1);
      for (size_type i = 0; i < M; ++i)
        if (std::abs(vec[i]) > R(1e-6) * nrm) {
          ph = -std::conj(vec[i]) / std::abs(vec[i]);
          break;
        }
      for (size_type i = 0; i < M; ++i) vec[i] = vec[i] * ph / nrm;
      // standardise: |E'_k> = sqrt(E_k / G_k) |E_k>,  G_k = <E_k| chi |E_k>   (chi_matrix.cpp:85-118)
      C Gk(0);
      for (size_type i = 0; i < M; ++i) {
        C Gi(0);
        for (size_type j = 0; j < M; ++j) Gi += this->data_[i][j] * vec[j];
        Gk += std::conj(vec[i]) * Gi;
      }
      assert(std::abs(std::imag(Gk)) < 1.0e-10 * (scale > 0 ? scale : 1) && "Error: rescale factor is not real.");
      if (std::abs(Gk) > 0) {
        const C f = std::sqrt(evalues_[k] / Gk);
        for (size_type i = 0; i < M; ++i) vec[i] *= f;
      }
    }
    NormalizeEigenProbAndRenormalizeEigenVect();
  }
  template <class U>
  void Solve(U *) {
    throw std::invalid_argument("ChiMatrix::SolveEigenSystem: only complex element types have an eigensystem");
  }

  void IdealHadamard(std::complex<double> *) {
    if (M != 4 || align != 32) {
      std::cout << "---- dummy version of EigensystemOfIdealHadamardChannel()\n";
      return;
    }
    for (size_type i = 0; i < M; ++i)
      for (size_type j = 0; j < M; ++j) {
        const bool xz = (i == 1 || i == 3) && (j == 1 || j == 3);
        assert(xz ? this->data_[i][j] == value_type(0.5, 0) : std::norm(this->data_[i][j]) == 0);
        (void)xz;
      }
    const value_type zero(0, 0), one(1, 0), h(1 / std::sqrt(2), 0);
    evalues_.assign(M, zero);
    evalues_[0] = one;
    evectors_.assign(M, std::vector<value_type>(M, zero));
    evectors_[0] = {zero, h, zero, h};
    evectors_[1] = {zero, h, zero, -h};
    evectors_[2] = {one, zero, zero, zero};
    evectors_[3] = {zero, zero, one, zero};
    NormalizeEigenProbAndRenormalizeEigenVect();
  }
  template <class U>
  void IdealHadamard(U *) {
    std::cout << "---- dummy version of EigensystemOfIdealHadamardChannel()\n";
  }
};

}  // namespace iqs
#endif
