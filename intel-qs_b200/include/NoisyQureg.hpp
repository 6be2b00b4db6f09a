// NoisyQureg.hpp -- register that inserts stochastic noise gates between the "experimental" gates
// of a circuit (interface and numerics of reference include/NoisyQureg.hpp:27-107, noise model
// :248-297).  Pure host logic over QubitRegister: every noise gate is one Apply1QubitGate on the
// device, so a noisy circuit costs (experimental + noise) sweeps.
//
// Model: every qubit carries the time elapsed since its last gate.  Before a gate on qubit q the
// idle period t is turned into a Pauli-twirled channel with p_X = p_Y = (1 - e^{-t/T1})/4,
// p_Z = (1 - e^{-t/T2})/2 + (1 - e^{-t/T1})/4, realised as the rotation
// exp(-i v_X X) exp(-i v_Y Y) exp(-i v_Z Z) with v_a = N(0,1) * sqrt(-log(1 - p_a)) / 2.
// The Gaussian numbers come from std::default_random_engine + std::normal_distribution seeded in
// the constructor, so a seeded run reproduces the reference's sequence of noise gates.
#ifndef NOISY_QUREG_HPP
#define NOISY_QUREG_HPP

#include <cmath>
#include <random>
#include <vector>

namespace iqs {

template <class Type = ComplexDP>
class NoisyQureg : public QubitRegister<Type> {
  typedef typename QubitRegister<Type>::BaseType BaseType;
  using Reg = QubitRegister<Type>;

 public:
  NoisyQureg(unsigned num_qubits, unsigned RNG_seed = 12345, BaseType T1 = 2000, BaseType T2 = 1000)
      : Reg(num_qubits), idle_(num_qubits, 0.), pair_counts_((std::size_t)num_qubits * num_qubits, 0u), T_1(T1), T_2(T2) {
    generator.seed(RNG_seed);
  }
  ~NoisyQureg() {}

  // noiseless (re)initialisation; gate counters and idle times restart
  void Initialize(std::string style, std::size_t base_index) {
    n_one_ = n_two_ = 0;
    ResetTimeForAllQubits();
    Reg::Initialize(style, base_index);
  }

  void ResetTimeForAllQubits() { std::fill(idle_.begin(), idle_.end(), BaseType(0)); }
  // noise for the idle time accumulated so far on every qubit (e.g. before the final measurement)
  void ApplyNoiseGatesOnAllQubits() {
    for (unsigned q = 0; q < this->num_qubits; ++q) NoiseGate(q);
    ResetTimeForAllQubits();
  }

  void SetDecoherenceTime(BaseType T1, BaseType T2) {
    T_1 = T1;
    T_2 = T2;
  }
  void SetGateDurations(BaseType one_qubit, BaseType two_qubit) {
    t_one_ = one_qubit;
    t_two_ = two_qubit;
  }

  unsigned GetTotalExperimentalGateCount() { return n_one_ + n_two_; }
  unsigned GetOneQubitExperimentalGateCount() { return n_one_; }
  unsigned GetTwoQubitExperimentalGateCount() { return n_two_; }
  // row q of the count matrix: [q] = one-qubit gates on q, [p] = two-qubit gates between q and p
  std::vector<unsigned> GetExperimentalGateCount(unsigned q) {
    const std::size_t n = this->num_qubits;
    return std::vector<unsigned>(pair_counts_.begin() + q * n, pair_counts_.begin() + (q + 1) * n);
  }
  unsigned GetExperimentalGateCount(unsigned q1, unsigned q2) { return pair_counts_[(std::size_t)q1 * this->num_qubits + q2]; }

  // bookkeeping of one experimental gate: noise for the idle period of its qubits, every clock
  // advanced by the gate duration (no gate parallelism), the clocks of its qubits restarted
  void AddNoiseOneQubitGate(unsigned const qubit) {
    NoiseGate(qubit);
    Advance(t_one_);
    idle_[qubit] = 0.;
    ++n_one_;
    ++pair_counts_[(std::size_t)qubit * this->num_qubits + qubit];
  }
  void AddNoiseTwoQubitGate(unsigned const q1, unsigned const q2) {
    NoiseGate(q1);
    NoiseGate(q2);
    Advance(t_two_);
    idle_[q1] = 0.;
    idle_[q2] = 0.;
    ++n_two_;
    ++pair_counts_[(std::size_t)q1 * this->num_qubits + q2];
    ++pair_counts_[(std::size_t)q2 * this->num_qubits + q1];
  }

  void NoiseGate(unsigned const qubit) {
    BaseType v_X, v_Y, v_Z;
    if (!DrawAngles(qubit, v_X, v_Y, v_Z)) return;
    // U = exp(-i v_X X) exp(-i v_Y Y) exp(-i v_Z Z) = [[A B, -A* C*], [A C, A* B*]] with
    // A = e^{-i v_Z}, B = cos v_X cos v_Y - i sin v_X sin v_Y, C = cos v_X sin v_Y - i sin v_X cos v_Y
    const Type A = {std::cos(v_Z), -std::sin(v_Z)};
    const Type B = {std::cos(v_X) * std::cos(v_Y), -std::sin(v_X) * std::sin(v_Y)};
    const Type C = {std::cos(v_X) * std::sin(v_Y), -std::sin(v_X) * std::cos(v_Y)};
    iqs::TinyMatrix<Type, 2, 2, 32> U_noise;
    U_noise(0, 0) = A * B;
    U_noise(0, 1) = -std::conj(A) * std::conj(C);
    U_noise(1, 0) = A * C;
    U_noise(1, 1) = std::conj(A) * std::conj(B);
    Reg::Apply1QubitGate(qubit, U_noise);
  }

  // Historical variant (reference :300-447, "should NOT be used"): the three small rotations are
  // composed as a 3x3 rotation matrix R_X(v_X) R_Y(v_Y) R_Z(v_Z), converted to axis-angle and
  // applied as exp(i angle/2 axis.sigma).
  void NoiseGate_OLD(unsigned const qubit) {
    BaseType v_X, v_Y, v_Z;
    if (!DrawAngles(qubit, v_X, v_Y, v_Z)) return;
    const BaseType cx = std::cos(v_X), sx = std::sin(v_X), cy = std::cos(v_Y), sy = std::sin(v_Y), cz = std::cos(v_Z), sz = std::sin(v_Z);
    // off-diagonal differences of R give 2 sin(angle) * axis
    const BaseType u[3] = {(cx * sy * sz + sx * cz) - (-sx * cy), sy - (-cx * sy * cz + sx * sz), (sx * sy * cz + cx * sz) - (-cy * sz)};
    const BaseType norm_u = std::sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
    const BaseType axis[3] = {u[0] / norm_u, u[1] / norm_u, u[2] / norm_u};
    const BaseType trace_R = cy * cz - sx * sy * sz + cx * cz + cx * cy;
    const BaseType angle = std::acos((trace_R - 1.) / 2.);
    const BaseType s = std::sin(angle / 2.), c = std::cos(angle / 2.);
    iqs::TinyMatrix<Type, 2, 2, 32> rot;
    rot(0, 0) = Type(c, s * axis[2]);
    rot(0, 1) = Type(s * axis[1], s * axis[0]);
    rot(1, 0) = Type(-s * axis[1], s * axis[0]);
    rot(1, 1) = Type(c, -s * axis[2]);
    Reg::Apply1QubitGate(qubit, rot);
  }

  // experimental gates: noise for the idle period first, then the gate itself
  void Apply1QubitGate(unsigned const q, iqs::TinyMatrix<Type, 2, 2, 32> V) {
    AddNoiseOneQubitGate(q);
    Reg::Apply1QubitGate(q, V);
  }
  void ApplyHadamard(unsigned const q) {
    AddNoiseOneQubitGate(q);
    Reg::ApplyHadamard(q);
  }
  void ApplyRotationX(unsigned const q, BaseType theta) {
    AddNoiseOneQubitGate(q);
    Reg::ApplyRotationX(q, theta);
  }
  void ApplyRotationY(unsigned const q, BaseType theta) {
    AddNoiseOneQubitGate(q);
    Reg::ApplyRotationY(q, theta);
  }
  void ApplyRotationZ(unsigned const q, BaseType theta) {
    AddNoiseOneQubitGate(q);
    Reg::ApplyRotationZ(q, theta);
  }
  void ApplyCPauliX(unsigned const q1, unsigned const q2) {
    AddNoiseTwoQubitGate(q1, q2);
    Reg::ApplyCPauliX(q1, q2);
  }
  void ApplyControlled1QubitGate(unsigned const q1, unsigned const q2, iqs::TinyMatrix<Type, 2, 2, 32> V) {
    AddNoiseTwoQubitGate(q1, q2);
    Reg::ApplyControlled1QubitGate(q1, q2, V);
  }

 private:
  void Advance(BaseType dt) {
    for (auto &t : idle_) t += dt;
  }
  // three Gaussian angles for the idle time of `qubit`; false when no time has passed
  bool DrawAngles(unsigned qubit, BaseType &v_X, BaseType &v_Y, BaseType &v_Z) {
    const BaseType t = idle_[qubit];
    if (t == 0) return false;
    const BaseType p_X = (1. - std::exp(-t / T_1)) / 4.;
    const BaseType p_Y = (1. - std::exp(-t / T_1)) / 4.;
    const BaseType p_Z = (1. - std::exp(-t / T_2)) / 2. + (1. - std::exp(-t / T_1)) / 4.;
    assert(p_X > 0 && p_Y > 0 && p_Z > 0);
    const BaseType s_X = std::sqrt(-std::log(1. - p_X));
    const BaseType s_Y = std::sqrt(-std::log(1. - p_Y));
    const BaseType s_Z = std::sqrt(-std::log(1. - p_Z));
    v_X = gaussian_RNG(generator) * s_X / 2.;
    v_Y = gaussian_RNG(generator) * s_Y / 2.;
    v_Z = gaussian_RNG(generator) * s_Z / 2.;
    return true;
  }

  std::vector<BaseType> idle_;          // time since the last gate, per qubit
  std::vector<unsigned> pair_counts_;   // num_qubits x num_qubits experimental gate counts
  unsigned n_one_ = 0, n_two_ = 0;
  BaseType t_one_ = 1., t_two_ = 1.;    // durations of experimental 1- and 2-qubit gates
  BaseType T_1 = 1000, T_2 = 100;
  std::default_random_engine generator;
  std::normal_distribution<BaseType> gaussian_RNG{0., 1.};
};

}  // namespace iqs
#endif
