// qureg_misc.cpp -- host-side I/O of the state (Print, ExportAmplitudes, dumpbin) and the noise
// front-ends.  These are clients of the hot path: they
// download the shard once and format it on the host, rank after rank.
// Output formats follow reference src/qureg_utils.cpp:455-689; noise gates src/qureg_noisysimul.cpp:15-91.
#include <cstdio>
#include <fstream>

#include "qureg_impl.hpp"

namespace iqs {

using detail::Check;

namespace {
template <class Type>
std::vector<Type> Snapshot(iqsb_state *dev, std::size_t n) {
  std::vector<Type> host(n);
  Check(iqsb_download(dev, host.data(), 0, n), "downloading the state");
  return host;
}
}  // namespace

template <class Type>
void QubitRegister<Type>::Print(std::string x, std::vector<std::size_t>) {
  FlushForRead();
  RestoreCanonicalPlacement();
  BeforeDeviceOp();
  int my_rank = iqs::mpi::Environment::GetStateRank(), nprocs = iqs::mpi::Environment::GetStateSize();
  std::vector<Type> host = Snapshot<Type>(dev_, LocalSize());
  double cumulative = 0;
  iqs::mpi::StateBarrier();
  for (int r = 0; r < nprocs; ++r) {
    if (r == my_rank) {
      if (r == 0) {
        printf("qubit permutation: %s\n", qubit_permutation->GetMapStr().c_str());
        printf("%s=[\n", x.c_str());
      }
      for (std::size_t i = 0; i < LocalSize(); ++i) {
        std::string bin = qubit_permutation->data2program((std::size_t)my_rank * LocalSize() + i);
        printf("\t%-13.8lf + i * %-13.8lf   %% |%s> p=%lf\n", (double)std::real(host[i]), (double)std::imag(host[i]), bin.c_str(),
               (double)std::norm(host[i]));
        cumulative += std::norm(host[i]);
      }
      fflush(stdout);
    }
    iqs::mpi::StateBarrier();
  }
  iqs::mpi::AllreduceDouble(&cumulative, 1, iqs::mpi::SUM);
  if (my_rank == 0) printf("]; %% cumulative probability = %lf\n", cumulative);
  iqs::mpi::StateBarrier();
}

template <class Type>
void QubitRegister<Type>::ExportAmplitudes(std::string ofname) {
  FlushForRead();
  RestoreCanonicalPlacement();
  BeforeDeviceOp();
  int my_rank = iqs::mpi::Environment::GetStateRank(), nprocs = iqs::mpi::Environment::GetStateSize();
  std::vector<Type> host = Snapshot<Type>(dev_, LocalSize());
  iqs::mpi::StateBarrier();
  for (int r = 0; r < nprocs; ++r) {
    if (r == my_rank) {
      std::ofstream of(ofname, std::ofstream::app);
      if (r == 0) of << "\t\"amplitudes\" :\n\t{" << std::endl;
      std::string str;
      char s[4096];
      for (std::size_t i = 0; i < LocalSize(); ++i) {
        std::string bin = qubit_permutation->data2program((std::size_t)my_rank * LocalSize() + i);
        snprintf(s, sizeof(s), "\t\t\"%s\" : [%-13.8lf, %-13.8lf, %lf],\n", bin.c_str(), (double)std::real(host[i]),
                 (double)std::imag(host[i]), (double)std::norm(host[i]));
        str += s;
      }
      if (r == nprocs - 1) str = str.substr(0, str.size() - 2) + "\n";  // no trailing comma
      of << str;
      if (r == nprocs - 1) of << "\t}\n";
    }
    iqs::mpi::StateBarrier();
  }
}

template <class Type>
void QubitRegister<Type>::dumpbin(std::string fn) {
  // raw amplitudes, rank r at byte offset r * LocalSize() * sizeof(Type) (the reference's MPI-IO layout)
  FlushForRead();
  RestoreCanonicalPlacement();
  BeforeDeviceOp();
  int my_rank = iqs::mpi::Environment::GetStateRank(), nprocs = iqs::mpi::Environment::GetStateSize();
  std::vector<Type> host = Snapshot<Type>(dev_, LocalSize());
  double t0 = sec();
  for (int r = 0; r < nprocs; ++r) {
    if (r == my_rank) {
      FILE *f = fopen(fn.c_str(), r == 0 ? "wb" : "r+b");
      if (!f) throw std::runtime_error("dumpbin: cannot open " + fn);
      fseek(f, (long)((std::size_t)r * LocalSize() * sizeof(Type)), SEEK_SET);
      fwrite(host.data(), sizeof(Type), LocalSize(), f);
      fclose(f);
    }
    iqs::mpi::StateBarrier();
  }
  double t1 = sec();
  if (my_rank == 0) printf("Dumping state to %s took %lf sec (%lf MB/s)\n", fn.c_str(), t1 - t0, double(sizeof(Type) * LocalSize()) / (t1 - t0) / 1e6);
}

// ---------------------------------------------------------------------------------------------
// noise: pure clients of Apply1QubitGate
// ---------------------------------------------------------------------------------------------
template <class Type>
void QubitRegister<Type>::SetNoiseTimescales(BaseType T1, BaseType T2) {
  assert(T2 >= T1 / 2.);
  T_1_ = T1;
  T_2_ = T2;
  T_phi_ = 1. / (1. / T2 - 1. / (2. * T1));
}

template <class Type>
void QubitRegister<Type>::ApplyNoiseGate(unsigned qubit, BaseType duration) {
  assert(rng_ptr_ != nullptr);
  if (duration == 0) return;
  // Pauli-twirl noise: U = exp(-i vX X) exp(-i vY Y) exp(-i vZ Z) with Gaussian angles whose
  // variances follow from T1 / T2 and the idle time
  BaseType decay1 = 1. - std::exp(-duration / T_1_), decay2 = 1. - std::exp(-duration / T_2_);
  BaseType p_X = decay1 / 4., p_Y = decay1 / 4., p_Z = decay2 / 2. + decay1 / 4.;
  assert(p_X > 0 && p_Y > 0 && p_Z > 0);
  BaseType s_X = std::sqrt(-std::log(1. - p_X)), s_Y = std::sqrt(-std::log(1. - p_Y)), s_Z = std::sqrt(-std::log(1. - p_Z));
  BaseType v_X, v_Y, v_Z;
  rng_ptr_->GaussianRandomNumbers(&v_X, 1, "state");
  v_X *= s_X / 2.;
  rng_ptr_->GaussianRandomNumbers(&v_Y, 1, "state");
  v_Y *= s_Y / 2.;
  rng_ptr_->GaussianRandomNumbers(&v_Z, 1, "state");
  v_Z *= s_Z / 2.;
  Type A = {std::cos(v_Z), -std::sin(v_Z)};
  Type B = {std::cos(v_X) * std::cos(v_Y), -std::sin(v_X) * std::sin(v_Y)};
  Type C = {std::cos(v_X) * std::sin(v_Y), -std::sin(v_X) * std::cos(v_Y)};
  TM2x2<Type> U_noise;
  U_noise(0, 0) = A * B;
  U_noise(0, 1) = -std::conj(A) * std::conj(C);
  U_noise(1, 0) = A * C;
  U_noise(1, 1) = std::conj(A) * std::conj(B);
  QubitRegister<Type>::Apply1QubitGate(qubit, U_noise);
}

// Quantum channels through the chi matrix (reference src/qureg_apply_channel.cpp:18-112): draw one
// eigen-operator of chi with probability |E_k| / sum |E_k| and apply it as a gate.  The operator is
// sum_i E_k,i sigma_i in the Pauli basis {id, X, Y, Z}; it is in general not unitary, averages over
// many trajectories reproduce the channel.
namespace {
template <class Type, class Chi>
unsigned DrawEigenOperator(iqs::RandomNumberGenerator<typename QubitRegister<Type>::BaseType> *rng, Chi &chi, unsigned dim) {
  typename QubitRegister<Type>::BaseType r;
  rng->UniformRandomNumbers(&r, 1, 0, 1, "state");
  unsigned k = 0;
  while (r > chi.GetEigenCumulativeProbability(k)) {
    ++k;
    if (k >= dim) {
      assert(0 && "Error: p_cum should be normalized to 1.");
      throw std::runtime_error("ApplyChannel: the eigen-probabilities of the chi matrix do not sum to 1 (was SolveEigenSystem called?)");
    }
  }
  return k;
}
}  // namespace

template <class Type>
void QubitRegister<Type>::ApplyChannel(const unsigned qubit, CM4x4<Type> &chi) {
  assert(rng_ptr_ != nullptr);
  const unsigned k = DrawEigenOperator<Type>(rng_ptr_, chi, 4);
  const std::vector<Type> e = chi.GetEigenVector(k);
  const Type I(0., 1.);
  TM2x2<Type> op;  // e0 id + e1 X + e2 Y + e3 Z
  op(0, 0) = e[0] + e[3];
  op(0, 1) = e[1] - I * e[2];
  op(1, 0) = e[1] + I * e[2];
  op(1, 1) = e[0] - e[3];
  if (std::real(chi.GetEigenValue(k)) < 0) overall_sign_of_channels *= -1;  // negative weight of the trajectory
  Apply1QubitGate(qubit, op);
}

template <class Type>
void QubitRegister<Type>::ApplyChannel(const unsigned qubit1, const unsigned qubit2, CM16x16<Type> &chi) {
  assert(rng_ptr_ != nullptr);
  const unsigned k = DrawEigenOperator<Type>(rng_ptr_, chi, 16);
  const std::vector<Type> e = chi.GetEigenVector(k);
  // op = sum_ab e[4a+b] sigma_a (x) sigma_b, rows/columns indexed 2*bit(qubit1) + bit(qubit2).  Like
  // the reference (apply_channel.cpp:87-104) the upper triangle and the diagonal are that expansion
  // and the lower triangle is its conjugate: the operator is taken as Hermitian.
  const Type I(0., 1.), one(1., 0.), zero(0., 0.);
  const Type pauli[4][2][2] = {{{one, zero}, {zero, one}}, {{zero, one}, {one, zero}}, {{zero, -I}, {I, zero}}, {{one, zero}, {zero, -one}}};
  TM4x4<Type> op;
  for (unsigned r = 0; r < 4; ++r)
    for (unsigned c = r; c < 4; ++c) {
      Type sum = zero;
      for (unsigned a = 0; a < 4; ++a)
        for (unsigned b = 0; b < 4; ++b) {
          const Type w = pauli[a][r >> 1][c >> 1] * pauli[b][r & 1][c & 1];
          if (w != zero) sum += w * e[4 * a + b];
        }
      op(r, c) = sum;
      if (c != r) op(c, r) = std::conj(sum);
    }
  Apply2QubitGate(qubit1, qubit2, op);
}

template class QubitRegister<ComplexSP>;
template class QubitRegister<ComplexDP>;

}  // namespace iqs
