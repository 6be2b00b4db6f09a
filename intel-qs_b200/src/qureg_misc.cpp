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

template <class Type>
void QubitRegister<Type>::ApplyChannel(const unsigned, CM4x4<Type> &) {
  throw std::runtime_error("QubitRegister::ApplyChannel: quantum channels (chi-matrix eigen-decomposition) are outside the scope of the B200 engine");
}
template <class Type>
void QubitRegister<Type>::ApplyChannel(const unsigned, const unsigned, CM16x16<Type> &) {
  throw std::runtime_error("QubitRegister::ApplyChannel: quantum channels (chi-matrix eigen-decomposition) are outside the scope of the B200 engine");
}

template class QubitRegister<ComplexSP>;
template class QubitRegister<ComplexDP>;

}  // namespace iqs
