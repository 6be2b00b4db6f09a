// chi_matrix_check.cpp -- host-only checks of iqs::ChiMatrix (no GPU, no library): the reference's
// known-answer test for the eigensystem (unit_test/include/chi_matrix_test.hpp:139-163), its
// container tests (:33-137), and reconstruction chi = sum_k E_k |E_k><E_k| / sum|E_k| for random
// Hermitian matrices of the sizes the channels use (4 and 16).  Prints "OK <name>" lines; any
// failure prints "FAILED" and exits 1.  Run by tests/test_chi_matrix.py.
#include <cstdio>
#include <random>

#include "chi_matrix.hpp"

using C = std::complex<double>;
static int failures = 0;
#define EXPECT(cond, name)                                   \
  do {                                                       \
    if (!(cond)) { printf("FAILED %s: %s\n", name, #cond); ++failures; } \
  } while (0)

template <unsigned N>
static void container() {
  iqs::ChiMatrix<C, N> mat;
  iqs::ChiMatrix<C, N, 32> mata;
  EXPECT(mat.numRows() == N && mat.numCols() == N && mat.size() == N * N, "size");
  for (unsigned i = 0; i < N; ++i)
    for (unsigned j = 0; j < N; ++j) {
      EXPECT(mat[i][j] == mat(i, j), "index");
      mat(i, j) = 1. + i + j;
    }
  iqs::ChiMatrix<C, N> const &matc(mat);
  mata = mat;
  iqs::ChiMatrix<C, N> matb = matc;
  iqs::ChiMatrix<C, N, 32> matd = matc;
  EXPECT(matb == mat && matb == mata && matb == matc && matc == matb && matd == mat && matd == mata && matc == matd, "copies");
  EXPECT(&mat(0, 0) == mat.getPtr(), "getPtr");
  mata(0, 0) = -100.;
  EXPECT(mata != matc && matc != mata, "not equal");
  printf("OK container<%u>\n", N);
}

static void assign() {
  double init[2][2] = {{1., 2.}, {3., 4.}};
  iqs::ChiMatrix<C, 2> mat = init;
  iqs::ChiMatrix<C, 2> mat2 = {{C(1.), C(2.)}, {C(3.), C(4.)}};
  for (unsigned i = 0; i < 2; ++i)
    for (unsigned j = 0; j < 2; ++j) EXPECT(mat(i, j) == 1. + 2. * i + j && mat2(i, j) == 1. + 2. * i + j, "assign");
  EXPECT(mat == init && mat2 == init && mat == mat2 && !(mat != init) && !(mat != mat2), "compare");
  mat = {{C(0.), C(1.)}, {C(1.), C(2.)}};
  EXPECT(mat(1, 1) == 2. && mat(0, 1) == 1., "reassign");
  printf("OK assign\n");
}

// chi_matrix_test.hpp:139-163: [[1,3],[3,7]] -> eigenvalues -0.24264069, 8.24264069; renormalised
// eigenvectors (-2.69121547, 1.11473795) and (-1.11473795, -2.69121547)
static void known_answer() {
  iqs::ChiMatrix<C, 2> chi = {{C(1.), C(3.)}, {C(3.), C(7.)}};
  chi.SolveEigenSystem();
  const double tol = 1e-7;
  EXPECT(std::abs(chi.GetEigenValue(0) - -0.24264069) < tol, "E0");
  EXPECT(std::abs(chi.GetEigenValue(1) - 8.24264069) < tol, "E1");
  EXPECT(std::abs(chi.GetEigenVector(0)[0] - -2.69121547) < tol && std::abs(chi.GetEigenVector(0)[1] - 1.11473795) < tol, "v0");
  EXPECT(std::abs(chi.GetEigenVector(1)[0] - -1.11473795) < tol && std::abs(chi.GetEigenVector(1)[1] - -2.69121547) < tol, "v1");
  EXPECT(std::abs(chi.GetEigenProbability(0) + chi.GetEigenProbability(1) - 1.) < 1e-15 && chi.GetEigenCumulativeProbability(1) == 1., "probabilities");
  printf("OK known_answer  E = %.8f %.8f  v0 = (%.8f, %.8f)\n", chi.GetEigenValue(0).real(), chi.GetEigenValue(1).real(), chi.GetEigenVector(0)[0].real(),
         chi.GetEigenVector(0)[1].real());
}

template <unsigned N>
static void reconstruct(unsigned seed) {
  std::mt19937_64 gen(seed);
  std::uniform_real_distribution<double> u(-1., 1.);
  iqs::ChiMatrix<C, N, 32> chi;
  for (unsigned i = 0; i < N; ++i) {
    chi(i, i) = C(u(gen), 0);
    for (unsigned j = i + 1; j < N; ++j) {
      chi(i, j) = C(u(gen), u(gen));
      chi(j, i) = std::conj(chi(i, j));
    }
  }
  chi.SolveEigenSystem();
  double total = 0, err = 0, ortho = 0, cum = 0;
  for (unsigned k = 0; k < N; ++k) {
    total += std::abs(chi.GetEigenValue(k).real());
    if (k) EXPECT(chi.GetEigenValue(k).real() >= chi.GetEigenValue(k - 1).real(), "ascending");
  }
  // stored vectors are sqrt(total) * unit vectors; sum_k sign(E_k) p_k-weighted projectors rebuild chi:
  // chi_ij = sum_k E_k u_k,i conj(u_k,j) = sum_k (E_k / total) w_k,i conj(w_k,j)
  for (unsigned i = 0; i < N; ++i)
    for (unsigned j = 0; j < N; ++j) {
      C sum(0);
      for (unsigned k = 0; k < N; ++k) sum += chi.GetEigenValue(k) / total * chi.GetEigenVector(k)[i] * std::conj(chi.GetEigenVector(k)[j]);
      err = std::max(err, std::abs(sum - chi(i, j)));
    }
  for (unsigned k = 0; k < N; ++k)
    for (unsigned l = 0; l < N; ++l) {
      C dot(0);
      for (unsigned i = 0; i < N; ++i) dot += std::conj(chi.GetEigenVector(k)[i]) * chi.GetEigenVector(l)[i];
      ortho = std::max(ortho, std::abs(dot / total - C(k == l ? 1. : 0.)));
    }
  for (unsigned k = 0; k < N; ++k) {
    cum += chi.GetEigenProbability(k);
    EXPECT(std::abs(cum - chi.GetEigenCumulativeProbability(k)) < 1e-14, "cumulative");
  }
  EXPECT(err < 1e-13 && ortho < 1e-13 && std::abs(cum - 1.) < 1e-14, "reconstruction");
  printf("OK reconstruct<%u> seed %u  max|chi - sum E|E><E|| = %.2e  orthonormality %.2e\n", N, seed, err, ortho);
}

static void hadamard_channel() {
  iqs::ChiMatrix<C, 4, 32> chi, closed;
  chi(1, 1) = chi(1, 3) = chi(3, 1) = chi(3, 3) = C(0.5, 0);
  closed = chi;
  chi.SolveEigenSystem();
  closed.EigensystemOfIdealHadamardChannel();
  // one eigenvalue 1 (last, ascending) with eigenvector +-(0,1,0,1)/sqrt(2); the rest has weight 0
  EXPECT(std::abs(chi.GetEigenValue(3).real() - 1.) < 1e-15 && std::abs(chi.GetEigenValue(0).real()) < 1e-15, "eigenvalues");
  EXPECT(chi.GetEigenCumulativeProbability(2) < 1e-15 && std::abs(chi.GetEigenCumulativeProbability(3) - 1.) < 1e-15, "probabilities");
  std::vector<C> v = chi.GetEigenVector(3), w = closed.GetEigenVector(0);
  for (unsigned i = 0; i < 4; ++i) EXPECT(std::abs(v[i] + w[i]) < 1e-15, "eigenvector (phase convention: first component negative)");
  EXPECT(closed.GetEigenValue(0) == C(1, 0) && closed.GetEigenProbability(0) == 1., "closed form");
  printf("OK hadamard_channel\n");
}

int main() {
  container<1>();
  container<2>();
  container<4>();
  assign();
  known_answer();
  for (unsigned s = 1; s <= 3; ++s) reconstruct<4>(s);
  for (unsigned s = 1; s <= 3; ++s) reconstruct<16>(s);
  hadamard_channel();
  bool threw = false;
  iqs::ChiMatrix<C, 2> bad = {{C(1.), C(3.)}, {C(2.), C(7.)}};
  try { bad.SolveEigenSystem(); } catch (std::invalid_argument const &) { threw = true; }
  EXPECT(threw, "non-Hermitian input is rejected");
  printf(failures ? "FAILED %d checks\n" : "ALL OK\n", failures);
  return failures ? 1 : 0;
}
