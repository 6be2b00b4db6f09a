// qureg_measure.cpp -- probabilities, collapse, classical-bit tests and expectation values.
//
// Reference behaviour restated: src/qureg_measure.cpp (IsClassicalBit :19-81, CollapseQubit :92-126,
// GetProbability :135-178, GetClassicalValue :183-262) and src/qureg_expectval.cpp (:18-375).
// The serial CPU sums become warp-shuffle reductions (csrc/kernels_reduce.cu) followed by an NCCL
// all-reduce of one double when the register spans several ranks.
#include <cstdlib>
#include <limits>

#include "qureg_impl.hpp"

namespace iqs {

using detail::Check;

template <class Type>
bool QubitRegister<Type>::IsClassicalBit(unsigned qubit, BaseType tolerance) const {
  assert(qubit < num_qubits);
  unsigned position = (*qubit_permutation)[qubit];
  assert(position < num_qubits);
  const_cast<QubitRegister *>(this)->FlushForRead();
  BeforeDeviceOp();
  std::size_t glb_start = UL(iqs::mpi::Environment::GetStateRank()) * LocalSize();
  int flags[2] = {0, 0};
  Check(iqsb_any_above(dev_, Phys(position), (double)tolerance, glb_start, flags), "IsClassicalBit");
  double v[2] = {double(flags[0]), double(flags[1])};
  iqs::mpi::AllreduceDouble(v, 2, iqs::mpi::MAX);  // logical OR over the ranks (measure.cpp:69-70)
  return !(v[0] > 0 && v[1] > 0);
}

template <class Type>
void QubitRegister<Type>::CollapseQubit(unsigned qubit, bool value) {
  assert(qubit < num_qubits);
  unsigned position = (*qubit_permutation)[qubit];
  assert(position < num_qubits);
  FlushForRead();
  BeforeDeviceOp();
  unsigned M = LocalQubits();
  const unsigned P = Phys(position);  // the physical bit that holds the position (src/placement.cpp)
  if (P < M) {
    Check(iqsb_collapse(dev_, P, value ? 1 : 0), "CollapseQubit");
  } else {
    std::size_t glb_start = UL(iqs::mpi::Environment::GetStateRank()) * LocalSize();
    if (check_bit(glb_start, P) != value) Check(iqsb_fill_const(dev_, 0., 0.), "CollapseQubit");
  }
}

template <class Type>
bool QubitRegister<Type>::OneSweepReductions() {
  static int on = -1;
  if (on < 0) {
    const char *e = getenv("IQS_B200_ONE_SWEEP");
    on = (e && *e == '0') ? 0 : 1;
  }
  return on != 0;
}

// The reference sums half of the state per call (src/qureg_measure.cpp:150-167).  Here the first
// call after a change does the same on the device (8 B per amplitude); a further call with nothing
// in between computes ALL marginals in one read (iqsb_prob_all, 16 B per amplitude) and the calls
// after that are free -- the 2 x n <Z> of a Heisenberg step cost 3 sweeps instead of 2 x n.
template <class Type>
typename QubitRegister<Type>::BaseType QubitRegister<Type>::GetProbability(unsigned qubit) {
  assert(qubit < num_qubits);
  unsigned position = (*qubit_permutation)[qubit];
  assert(position < num_qubits);
  const unsigned M = LocalQubits(), n = (unsigned)num_qubits;
  const std::size_t glb_start = UL(iqs::mpi::Environment::GetStateRank()) * LocalSize();
  const bool cacheable = OneSweepReductions() && !raw_exposed_ && M >= 1 && M <= 38 && n <= 64;
  const bool quiet = queue_.empty() && checked_out_.empty() && !host_touched_;
  if (cacheable && quiet && marginal_queries_ > 0 && marginals_.size() == n) {
    const unsigned P = Phys(position);
    if (marginals_[P] == marginals_[P]) return (BaseType)marginals_[P];  // known (not NaN)
    if (!marginals_all_) {
      double loc[40];
      Check(iqsb_prob_all(dev_, loc, (int)M + 1), "GetProbability (all marginals)");
      std::vector<double> v(n, 0.);
      for (unsigned b = 0; b < n; ++b) v[b] = b < M ? loc[1 + b] : (check_bit(glb_start, b) ? loc[0] : 0.);
      iqs::mpi::AllreduceDouble(v.data(), (int)n, iqs::mpi::SUM);
      for (unsigned b = 0; b < n; ++b)
        if (!(marginals_[b] == marginals_[b])) marginals_[b] = v[b];  // keep what earlier calls returned
      marginals_all_ = true;
    }
    return (BaseType)marginals_[P];
  }
  FlushForRead();
  BeforeDeviceOp();
  const unsigned P = Phys(position);
  double p = 0.;
  if (P < M) {
    Check(iqsb_prob1(dev_, P, &p), "GetProbability");
  } else {
    if (check_bit(glb_start, P) == 1) Check(iqsb_norm2(dev_, &p), "GetProbability");
  }
  iqs::mpi::AllreduceDouble(&p, 1, iqs::mpi::SUM);
  if (cacheable) {
    marginals_.assign(n, std::numeric_limits<double>::quiet_NaN());
    marginals_[P] = p;
    marginals_all_ = false;
    marginal_queries_ = 1;
  }
  return (BaseType)p;
}

// <psi| P |psi> for a Pauli string whose X / Y factors sit on local physical bits: one read of the
// state, nothing written (iqsb_pauli_expect).  Returns false when the string needs the reference's
// basis-change sweeps instead (an X / Y factor on a rank bit, or IQS_B200_ONE_SWEEP=0).
template <class Type>
bool QubitRegister<Type>::PauliStringReadOnly(const std::vector<unsigned> &qubits, const std::vector<unsigned> &observables, double &value,
                                              double *norm2) {
  if (!OneSweepReductions() || LocalSize() < 2) return false;
  if (timer != nullptr) return false;  // EnableStatistics(): the per-gate timing lines of the basis changes are part of the output
  FlushForRead();
  const unsigned M = LocalQubits();
  uint64_t mask[4] = {0, 0, 0, 0};
  for (std::size_t i = 0; i < qubits.size(); ++i) {
    const uint64_t bit = uint64_t(1) << Phys((*qubit_permutation)[qubits[i]]);
    if ((mask[1] | mask[2] | mask[3]) & bit) return false;  // a qubit named twice: the sweeps define what that means
    mask[observables[i]] |= bit;
  }
  if (((mask[1] | mask[2]) >> M) != 0) return false;
  const bool quiet = checked_out_.empty() && !host_touched_;
  if (!quiet) BeforeDeviceOp();  // host writes go back first (that is a change: marginals are dropped)
  const std::size_t glb_start = UL(iqs::mpi::Environment::GetStateRank()) * LocalSize();
  double v[2] = {0., 0.};
  Check(iqsb_pauli_expect(dev_, mask[1], mask[2], mask[3], glb_start, v), "ExpectationValue");
  iqs::mpi::AllreduceDouble(v, 2, iqs::mpi::SUM);
  value = v[0];
  if (norm2) *norm2 = v[1];
  // the reference counts the two basis-change gates of every X / Y factor (ApplyHadamard / Apply1QubitGate
  // before and after the sum, src/qureg_expectval.cpp:148-210): programs that print the counter see the same numbers
  if (gate_counter != nullptr)
    for (std::size_t i = 0; i < qubits.size(); ++i)
      if (observables[i] == 1 || observables[i] == 2) {
        gate_counter->OneQubitIncrement(qubits[i]);
        gate_counter->OneQubitIncrement(qubits[i]);
      }
  return true;
}

template <class Type>
bool QubitRegister<Type>::GetClassicalValue(unsigned qubit, BaseType tolerance) const {
  assert(qubit < num_qubits);
  unsigned position = (*qubit_permutation)[qubit];
  assert(position < num_qubits);
  const_cast<QubitRegister *>(this)->FlushForRead();
  BeforeDeviceOp();
  std::size_t glb_start = UL(iqs::mpi::Environment::GetStateRank()) * LocalSize();
  int flags[2] = {0, 0};
  Check(iqsb_any_above(dev_, Phys(position), (double)tolerance, glb_start, flags), "GetClassicalValue");
  double v[2] = {double(flags[0]), double(flags[1])};
  iqs::mpi::AllreduceDouble(v, 2, iqs::mpi::MAX);
  bool zero = v[0] > 0, one = v[1] > 0;
  if (zero && !one) return false;
  if (!zero && one) return true;
  assert(false && "GetClassicalValue: the qubit is not in a classical state");
  return false;
}

// ---------------------------------------------------------------------------------------------
// expectation values
// ---------------------------------------------------------------------------------------------
template <class Type>
typename QubitRegister<Type>::BaseType QubitRegister<Type>::ExpectationValueX(unsigned qubit, BaseType coeff) {
  // the reference returns 1 - 2 P(1 after H) = <X> + (1 - <psi|psi>): the same number from one read
  double v = 0., n2 = 1.;
  if (PauliStringReadOnly({qubit}, {1u}, v, &n2)) return coeff * (BaseType)(v + (1. - n2));
  // <X> = <psi| H.Z.H |psi>
  ApplyHadamard(qubit);
  BaseType expectation = 1. - 2. * GetProbability(qubit);
  ApplyHadamard(qubit);
  return coeff * expectation;
}

template <class Type>
typename QubitRegister<Type>::BaseType QubitRegister<Type>::ExpectationValueY(unsigned qubit, BaseType coeff) {
  double v = 0., n2 = 1.;
  if (PauliStringReadOnly({qubit}, {2u}, v, &n2)) return coeff * (BaseType)(v + (1. - n2));
  // G^dagger.Z.G = Y
  TM2x2<Type> G;
  BaseType f = 1. / std::sqrt(2.);
  G(0, 0) = G(1, 0) = Type(f, 0.);
  G(0, 1) = Type(0., -f);
  G(1, 1) = Type(0., f);
  Apply1QubitGate(qubit, G);
  BaseType expectation = 1. - 2. * GetProbability(qubit);
  G(0, 0) = G(0, 1) = Type(f, 0.);
  G(1, 0) = Type(0., f);
  G(1, 1) = Type(0., -f);
  Apply1QubitGate(qubit, G);
  return coeff * expectation;
}

template <class Type>
typename QubitRegister<Type>::BaseType QubitRegister<Type>::ExpectationValueZ(unsigned qubit, BaseType coeff) {
  BaseType expectation = 1. - 2. * GetProbability(qubit);
  return coeff * expectation;
}

// observable: 1 == PauliX, 2 == PauliY, 3 == PauliZ
template <class Type>
typename QubitRegister<Type>::BaseType QubitRegister<Type>::ExpectationValue(std::vector<unsigned> &qubits,
                                                                            std::vector<unsigned> &observables, BaseType coeff) {
  assert(qubits.size() == observables.size());
  for (unsigned j = 0; j < qubits.size(); ++j) {
    assert(qubits[j] < num_qubits);
    assert(observables[j] > 0 && observables[j] < 4);
  }
  if (qubits.size() == 0) return coeff;
  if (qubits.size() == 1) {
    if (observables[0] == 1) return ExpectationValueX(qubits[0], coeff);
    if (observables[0] == 2) return ExpectationValueY(qubits[0], coeff);
    if (observables[0] == 3) return ExpectationValueZ(qubits[0], coeff);
  }
  {
    double v = 0.;
    if (PauliStringReadOnly(qubits, observables, v, nullptr)) return coeff * (BaseType)v;
  }
  TM2x2<Type> G, Ginv;
  BaseType f = 1. / std::sqrt(2.);
  G(0, 0) = G(1, 0) = Type(f, 0.);
  G(0, 1) = Type(0., -f);
  G(1, 1) = Type(0., f);
  Ginv(0, 0) = Ginv(0, 1) = Type(f, 0.);
  Ginv(1, 0) = Type(0., f);
  Ginv(1, 1) = Type(0., -f);

  for (std::size_t i = 0; i < qubits.size(); i++) {
    if (observables[i] == 1) ApplyHadamard(qubits[i]);
    else if (observables[i] == 2) Apply1QubitGate(qubits[i], G);
  }

  // signed sum of |a|^2, sign = parity of the involved bits of the GLOBAL index.  The mask is
  // built with 64-bit shifts (the reference's `1 << position` overflows at position 31, :170).
  std::size_t myrank = iqs::mpi::Environment::GetStateRank();
  std::size_t glb_start = UL(myrank) * LocalSize();
  std::size_t y = 0;
  FlushForRead();  // may move qubits between local and rank bits: build the mask afterwards
  for (std::size_t i = 0; i < qubits.size(); i++) y += std::size_t(1) << Phys((*qubit_permutation)[qubits[i]]);
  BeforeDeviceOp();
  double v = 0;
  Check(iqsb_parity_expect(dev_, y, glb_start, &v), "ExpectationValue");
  iqs::mpi::AllreduceDouble(&v, 1, iqs::mpi::SUM);
  BaseType expectation = (BaseType)v;

  for (std::size_t i = 0; i < qubits.size(); i++) {
    if (observables[i] == 1) ApplyHadamard(qubits[i]);
    else if (observables[i] == 2) Apply1QubitGate(qubits[i], Ginv);
  }
  return coeff * expectation;
}

#define IQS_EXPECT2(NAME, O1, O2)                                                                                   \
  template <class Type>                                                                                             \
  typename QubitRegister<Type>::BaseType QubitRegister<Type>::NAME(unsigned qubit, unsigned qubit2, BaseType coeff) { \
    std::vector<unsigned> qubits = {qubit, qubit2};                                                                 \
    std::vector<unsigned> observables = {O1, O2};                                                                   \
    return this->ExpectationValue(qubits, observables, coeff);                                                      \
  }
IQS_EXPECT2(ExpectationValueXX, 1, 1)
IQS_EXPECT2(ExpectationValueXY, 1, 2)
IQS_EXPECT2(ExpectationValueXZ, 1, 3)
IQS_EXPECT2(ExpectationValueYX, 2, 1)
IQS_EXPECT2(ExpectationValueYY, 2, 2)
IQS_EXPECT2(ExpectationValueYZ, 2, 3)
IQS_EXPECT2(ExpectationValueZX, 3, 1)
IQS_EXPECT2(ExpectationValueZY, 3, 2)
IQS_EXPECT2(ExpectationValueZZ, 3, 3)
#undef IQS_EXPECT2

template class QubitRegister<ComplexSP>;
template class QubitRegister<ComplexDP>;

}  // namespace iqs
