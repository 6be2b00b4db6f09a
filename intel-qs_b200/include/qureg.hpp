// qureg.hpp -- iqs::QubitRegister<Type>: the Intel-QS state-vector class, B200 edition.
//
// Public surface = reference include/qureg.hpp:127-417 (every public method and data member keeps
// its name, signature and meaning) so that programs written against Intel-QS compile unchanged.
// What is different is underneath: the 2^n amplitudes live in the HBM of one B200 per rank and
// every method is thin host code over the C ABI of libiqs_b200.so (include/iqsb.h), which
// launches hand-written sm_100a kernels.  There is no CPU fallback.
//
// Host view of the state (`operator[]`, `RawState()`, public `state`):
//  * one rank (default): the shard is CUDA managed memory whose preferred location is the GPU, so
//    `state` is a real host pointer; `operator[]`/`RawState()` first synchronise the engine's
//    stream, and the next gate prefetches the touched pages back to HBM.
//  * several ranks (or IQS_B200_MEM=device): the shard is plain device memory (it must be
//    cudaIpc-exportable for the NVLink peer kernels); `operator[]` then works on a host mirror
//    that is checked out chunk-wise on access and written back before the next device operation.
#pragma once

#include <algorithm>
#include <cassert>
#include <cmath>
#include <fstream>
#include <functional>
#include <iostream>
#include <limits>
#include <map>
#include <memory>
#include <numeric>
#include <string>
#include <tuple>
#include <vector>

#include "alignedallocator.hpp"
#include "bitops.hpp"
#include "chi_matrix.hpp"
#include "conversion.hpp"
#include "gate_counter.hpp"
#include "gate_spec.hpp"
#include "mpi_env.hpp"
#include "mpi_utils.hpp"
#include "permutation.hpp"
#include "rng_utils.hpp"
#include "timer.hpp"
#include "tinymatrix.hpp"
#include "utils.hpp"

template <class Type>
using TM2x2 = iqs::TinyMatrix<Type, 2, 2, 32>;
template <class Type>
using TM4x4 = iqs::TinyMatrix<Type, 4, 4, 32>;
template <class Type>
using CM4x4 = iqs::ChiMatrix<Type, 4, 32>;
template <class Type>
using CM16x16 = iqs::ChiMatrix<Type, 16, 32>;

struct iqsb_state;

namespace iqs {

// 'position' = data qubit (bit of the amplitude index), 'qubit' = program qubit; the
// qubit_permutation maps one to the other (reference qureg.hpp:80-125).
template <class Type = ComplexDP>
class QubitRegister {
 public:
  using value_type = Type;
  typedef typename extract_value_type<Type>::value_type BaseType;

  // Constructors and destructor.
  QubitRegister();
  QubitRegister(std::size_t num_qubits, std::string style = "", std::size_t base_index = 0,
                std::size_t tmp_spacesize_ = 0);
  QubitRegister(const QubitRegister &in);
  // Imported state: `state` is a caller-owned HOST buffer of LocalSize() amplitudes.  It is uploaded
  // here and serves as the register's host mirror: it is refreshed by SyncToHost(), operator[] and
  // on destruction (the reference computes in place in that buffer, qureg_init.cpp:193-200).
  QubitRegister(std::size_t num_qubits, Type *state, std::size_t tmp_spacesize_ = 0);
  ~QubitRegister();

  // Allocation & initialization
  void AllocateAdditionalQubit();
  void Allocate(std::size_t new_num_qubits, std::size_t tmp_spacesize_);
  void Initialize(std::size_t new_num_qubits, std::size_t tmp_spacesize_);
  void Initialize(std::string style, std::size_t base_index);

  // Amplitude at a local index (host reference; see the header comment).
  inline Type &operator[](std::size_t index) { return *HostAmplitude(index); }
  inline Type &operator[](std::size_t index) const { return *HostAmplitude(index); }
  Type GetGlobalAmplitude(std::size_t index) const;
  void SetGlobalAmplitude(std::size_t index, Type value);

  std::size_t LocalSize() const { return local_size_; }
  std::size_t GlobalSize() const { return global_size_; }

  void Resize(std::size_t new_num_amplitudes);
  std::size_t size() const { return global_size_; }
  std::size_t NumQubits() const { return num_qubits; }
  Type *TmpSpace() const { return state + LocalSize(); }
  size_t TmpSize() const { return tmp_spacesize_; }

  // Host pointer to the whole shard (NumPy buffer protocol).
  Type *RawState();

  inline bool check_bit(std::size_t variable, std::size_t position) const { return (variable >> position) & std::size_t(1); }
  inline std::size_t set_bit(std::size_t variable, std::size_t position) const { return variable | (std::size_t(1) << position); }
  inline std::size_t clear_bit(std::size_t variable, std::size_t position) const { return variable & ~(std::size_t(1) << position); }

  void EnableStatistics();
  void GetStatistics();
  void DisableStatistics();
  void ResetStatistics();

  // Permutation of the qubit order
  void PermuteQubits(std::vector<std::size_t> new_map, std::string style_of_map = "direct");
  void PermuteLocalQubits(std::vector<std::size_t> new_map, std::string style_of_map = "direct");
  void PermuteGlobalQubits(std::vector<std::size_t> new_map, std::string style_of_map = "direct");
  void PermuteByLocalGlobalExchangeOfQubitPairs(std::vector<std::size_t> new_map, std::string style_of_map = "direct");
  void EmulateSwap(unsigned qubit1, unsigned qubit2);

  // Generic gates
  bool Apply1QubitGate_helper(unsigned qubit, TM2x2<Type> const &m, std::size_t sstate_ind, std::size_t estate_ind,
                              iqs::GateSpec1Q spec = iqs::GateSpec1Q::None, BaseType angle = 0);
  void Apply1QubitGate(unsigned qubit, TM2x2<Type> const &m, iqs::GateSpec1Q spec = iqs::GateSpec1Q::None, BaseType angle = 0);
  bool ApplyControlled1QubitGate_helper(unsigned control_qubit, unsigned target_qubit, TM2x2<Type> const &m,
                                        std::size_t sind, std::size_t eind, iqs::GateSpec2Q spec = iqs::GateSpec2Q::None,
                                        BaseType angle = 0);
  void ApplyControlled1QubitGate(unsigned control_qubit, unsigned target_qubit, TM2x2<Type> const &m,
                                 iqs::GateSpec2Q spec = iqs::GateSpec2Q::None, BaseType angle = 0);
  // swap gates
  bool ApplySwap_helper(unsigned qubit1, unsigned qubit2, TM2x2<Type> const &m);
  void ApplySwap(unsigned qubit1, unsigned qubit2);
  void ApplyISwap(unsigned qubit1, unsigned qubit2);
  void Apply4thRootISwap(unsigned qubit1, unsigned qubit2);
  void ApplySqrtISwap(unsigned qubit1, unsigned qubit2);
  void ApplyISwapRotation(unsigned qubit1, unsigned qubit2, TM2x2<Type> const &m);
  // diagonal gates
  void ApplyDiagSimp(unsigned qubit1, unsigned qubit2, TM4x4<Type> const &m);
  void ApplyDiag(unsigned qubit1, unsigned qubit2, TM4x4<Type> const &m);
  void ApplyDiagControl(unsigned qubit1, unsigned qubit2, TM4x4<Type> const &m);
  void ApplyDiagGeneral(unsigned qubit1, unsigned qubit2, TM4x4<Type> const &m);
  // two-qubit gates
  void Apply2QubitGate(unsigned const qubit_high, unsigned const qubit_low, TM4x4<Type> const &m);
  // specialized gates
  void ApplyRotationX(unsigned const qubit, BaseType theta);
  void ApplyRotationY(unsigned const qubit, BaseType theta);
  void ApplyRotationZ(unsigned const qubit, BaseType theta);
  void ApplyPauliX(unsigned const qubit);
  void ApplyPauliY(unsigned const qubit);
  void ApplyPauliZ(unsigned const qubit);
  void ApplyPauliSqrtX(unsigned const qubit);
  void ApplyPauliSqrtY(unsigned const qubit);
  void ApplyPauliSqrtZ(unsigned const qubit);
  void ApplyT(unsigned const qubit);
  void ApplyToffoli(unsigned const qubit1, unsigned const qubit2, unsigned const qubit3);
  void ApplyHadamard(unsigned const qubit);
  void ApplyRotationXY(unsigned const qubit, BaseType phi, BaseType theta);

  void ApplyCRotationX(unsigned const control_qubit, unsigned const target_qubit, BaseType theta);
  void ApplyCRotationY(unsigned const control_qubit, unsigned const target_qubit, BaseType theta);
  void ApplyCRotationZ(unsigned const control_qubit, unsigned const target_qubit, BaseType theta);
  void ApplyCPauliX(unsigned const control_qubit, unsigned const target_qubit);
  void ApplyCPauliY(unsigned const control_qubit, unsigned const target_qubit);
  void ApplyCPauliZ(unsigned const control_qubit, unsigned const target_qubit);
  void ApplyCPauliSqrtZ(unsigned const control_qubit, unsigned const target_qubit);
  void ApplyCHadamard(unsigned const control_qubit, unsigned const target_qubit);
  void ApplyCPhaseRotation(unsigned const qubit, unsigned const qubit2, BaseType theta);

  // fusion
  void TurnOnFusion(unsigned log2llc = 20);
  void TurnOffFusion();
  bool IsFusionEnabled();
  void ApplyFusedGates();

  // gate specialization: hints only on the GPU (results are the same in every mode)
  void TurnOnSpecialize();
  void TurnOffSpecialize();
  void TurnOnSpecializeV2();
  void TurnOffSpecializeV2();

  // measurement
  bool GetClassicalValue(unsigned qubit, BaseType tolerance = 1.e-13) const;
  bool IsClassicalBit(unsigned qubit, BaseType tolerance = 1.e-13) const;
  void CollapseQubit(unsigned qubit, bool value);
  BaseType GetProbability(unsigned qubit);

  // expectation values without state update
  BaseType ExpectationValueX(unsigned const qubit, BaseType coeff = 1.);
  BaseType ExpectationValueY(unsigned const qubit, BaseType coeff = 1.);
  BaseType ExpectationValueZ(unsigned const qubit, BaseType coeff = 1.);
  BaseType ExpectationValueXX(unsigned const qubit, unsigned const qubit2, BaseType coeff = 1.);
  BaseType ExpectationValueXY(unsigned const qubit, unsigned const qubit2, BaseType coeff = 1.);
  BaseType ExpectationValueXZ(unsigned const qubit, unsigned const qubit2, BaseType coeff = 1.);
  BaseType ExpectationValueYX(unsigned const qubit, unsigned const qubit2, BaseType coeff = 1.);
  BaseType ExpectationValueYY(unsigned const qubit, unsigned const qubit2, BaseType coeff = 1.);
  BaseType ExpectationValueYZ(unsigned const qubit, unsigned const qubit2, BaseType coeff = 1.);
  BaseType ExpectationValueZX(unsigned const qubit, unsigned const qubit2, BaseType coeff = 1.);
  BaseType ExpectationValueZY(unsigned const qubit, unsigned const qubit2, BaseType coeff = 1.);
  BaseType ExpectationValueZZ(unsigned const qubit, unsigned const qubit2, BaseType coeff = 1.);
  BaseType ExpectationValue(std::vector<unsigned> &qubits, std::vector<unsigned> &observables, BaseType coeff = 1.);

  // noisy simulation (clients of the gate path; channels are outside the B200 scope)
  BaseType GetT1() { return T_1_; }
  BaseType GetT2() { return T_2_; }
  BaseType GetTphi() { return T_phi_; }
  void SetNoiseTimescales(BaseType T1, BaseType T2);
  void ApplyNoiseGate(const unsigned qubit, const BaseType duration);
  void ApplyChannel(const unsigned qubit, CM4x4<Type> &chi);
  void ApplyChannel(const unsigned qubit1, const unsigned qubit2, CM16x16<Type> &chi);
  BaseType GetOverallSignOfChannels() { return overall_sign_of_channels; }

  // Utilities
  bool operator==(const QubitRegister &rhs);
  BaseType MaxAbsDiff(QubitRegister &x, Type sfactor = Type(1, 0));
  BaseType MaxL2NormDiff(QubitRegister &x);
  void dumpbin(std::string fn);
  double Entropy();
  std::vector<double> GoogleStats();
  void Normalize();
  BaseType ComputeNorm();
  void InitializationWithSameAmplitudeEverywhere(Type amplitude);
  void AmplitudeWiseScalarMultiplication(Type factor);
  void AmplitudeWiseSum(QubitRegister<Type> &psi, Type factor = Type(1, 0));
  Type ComputeOverlap(QubitRegister<Type> &psi);

  void Print(std::string x, std::vector<std::size_t> qbits = {});
  void ExportAmplitudes(std::string ofname);

  // gates on global qubits: one fused compute+exchange kernel per rank over NVLink peer memory
  double HP_Distrpair(unsigned position, TM2x2<Type> const &m, iqs::GateSpec1Q spec = iqs::GateSpec1Q::None, BaseType angle = 0);
  double HP_Distrpair(unsigned control_position, unsigned target_position, TM2x2<Type> const &m,
                      iqs::GateSpec2Q spec = iqs::GateSpec2Q::None, BaseType angle = 0);
  double HP_DistrSwap(unsigned low_position, unsigned high_position, TM2x2<Type> const &m);

  // random number generator
  iqs::RandomNumberGenerator<BaseType> *GetRngPtr() { return rng_ptr_; }
  void ResetRngPtr() { rng_ptr_ = nullptr; }
  void SetRngPtr(iqs::RandomNumberGenerator<BaseType> *rng_ptr) { rng_ptr_ = rng_ptr; }
  void SetSeedRngPtr(std::size_t seed) {
    assert(rng_ptr_);
    rng_ptr_->SetSeedStreamPtrs(seed);
  }

  // ---- B200 extensions (not in the reference) ----------------------------------------------
  // complete all queued device work (and fused gates) and refresh the host mirror, if any
  void SyncToHost();
  // device handle of the shard (C ABI, include/iqsb.h)
  iqsb_state *DeviceState() { return dev_; }
  // make the device copy current and in the reference's amplitude order: run queued gates, undo
  // the placement layer's qubit moves, write back host-side edits
  void PrepareDevice() const {
    const_cast<QubitRegister *>(this)->FlushForRead();
    RestoreCanonicalPlacement();
    BeforeDeviceOp();
  }
  iqsb_state *DeviceState() const { return dev_; }
  // placement-layer statistics: multi-bit exchanges run so far and qubits moved by them
  void GetPlacementStatistics(std::size_t &exchanges, std::size_t &moved_qubits) const {
    exchanges = exchanges_;
    moved_qubits = exchanged_bits_;
  }

  // Members (public in the reference)
  std::size_t num_qubits;
  std::vector<Type, iqs::AlignedAllocator<Type, 256>> state_storage;  // unused: the state lives in HBM
  Type *state;  // host view of the shard (see the header comment)
  Permutation *qubit_permutation;
  Timer *timer;
  GateCounter *gate_counter;
  std::size_t llc_watermarkbit;
  bool imported_state;
  bool specialize;
  bool specialize2 = false;
  BaseType overall_sign_of_channels = 1;

  // fusion window (the reference's queue type, include/qureg.hpp:398; kept for source compatibility:
  // this engine queues gates, already resolved to positions, in the private `queue_` below)
  bool fusion;
  unsigned log2llc;
  std::vector<std::tuple<std::string, TM2x2<Type>, unsigned, unsigned>> fwindow;

  static void SetDoPrintExtraInfo(bool value) { do_print_extra_info = value; }

 private:
  std::size_t local_size_ = 0;
  std::size_t global_size_ = 0;
  std::size_t tmp_spacesize_ = 0;
  static bool do_print_extra_info;

  iqs::RandomNumberGenerator<BaseType> *rng_ptr_ = nullptr;
  BaseType T_1_ = 0, T_2_ = 0, T_phi_ = 0;

  // ---- device side -----------------------------------------------------------------------
  iqsb_state *dev_ = nullptr;
  bool managed_ = false;               // state is managed memory (host pointer valid)
  mutable bool host_touched_ = false;  // managed pages may sit on the host: prefetch before the next kernel
  // host mirror for device-memory registers
  mutable Type *mirror_ = nullptr;
  mutable bool mirror_owned_ = false;
  mutable std::vector<std::size_t> checked_out_;  // chunks downloaded since the last device op
  mutable std::vector<unsigned char> chunk_present_;

  // ---- gate queue and placement layer (src/placement.cpp) ----------------------------------
  // Gates wait in `queue_` (positions, program order) while fusion is on, and -- when the register
  // spans several GPUs -- for a short look-ahead window also while it is off: the scheduler uses
  // the gates it can see to decide which qubits to hold in the rank bits.  `place_[position]` is the
  // PHYSICAL bit of the distributed index that currently holds a position (identity = the
  // reference's layout: bits >= LocalQubits() are rank bits).  Gates are issued on physical bits;
  // anything that exposes the raw amplitude order first restores the identity placement.
  struct QueuedGate {
    int kind;  // 0: 1-qubit gate, 1: controlled gate
    unsigned control, target;  // positions
    bool diagonal;
    double m[8];
  };
  mutable std::vector<QueuedGate> queue_;
  mutable std::vector<uint8_t> place_, where_;  // position -> physical bit, physical bit -> position
  mutable std::vector<uint64_t> last_use_;
  mutable uint64_t use_clock_ = 0;
  mutable bool moved_ = false;  // some position may sit on a physical bit other than its own
  bool placement_ = false;   // several ranks and IQS_B200_PLACEMENT != 0
  unsigned lookahead_ = 0;   // unfused gates deferred for placement decisions (0: execute at once)
  mutable uint64_t exchanges_ = 0, exchanged_bits_ = 0;

  // ---- one-sweep reductions (src/qureg_measure.cpp) -------------------------------------------
  // P(bit = 1) of every physical bit, NaN = not known; valid until the next device operation or host
  // access.  The first GetProbability after a change reads half the state (iqsb_prob1); a second
  // one with nothing in between computes all marginals in one read (iqsb_prob_all).
  mutable std::vector<double> marginals_;
  mutable bool marginals_all_ = false;
  mutable unsigned marginal_queries_ = 0;
  mutable bool raw_exposed_ = false;  // RawState() handed the managed pointer out: no caching any more
  void InvalidateMarginals() const {
    marginals_all_ = false;
    marginal_queries_ = 0;
  }
  static bool OneSweepReductions();  // IQS_B200_ONE_SWEEP != 0 (default on)
  bool PauliStringReadOnly(const std::vector<unsigned> &qubits, const std::vector<unsigned> &observables, double &value, double *norm2);

  void InitPlacement();
  unsigned Phys(unsigned position) const { return place_.empty() ? position : place_[position]; }
  bool CanonicalPlacement() const { return !moved_; }
  bool Deferring() const { return fusion || (lookahead_ > 0 && timer == nullptr); }
  void Enqueue(int kind, unsigned control_position, unsigned target_position, TM2x2<Type> const &m);
  void RunQueue(std::size_t count);
  void ExecFusedRange(std::size_t first, std::size_t last);
  void ExecGate1(unsigned position, const double mm[8], bool diagonal, std::size_t sind, std::size_t eind, const std::string &name);
  bool ExecCGate1(unsigned control_position, unsigned target_position, const double mm[8], bool diagonal, std::size_t sind,
                  std::size_t eind, const std::string &name, TM2x2<Type> const *m);
  void BringLocal(std::size_t queue_from, uint64_t protect_mask);
  void SwapPlacement(unsigned position_a, unsigned position_b);
  void RestoreCanonicalPlacement() const;
  void Relabel(const Permutation &target);  // src/qureg_permute.cpp: a permutation is a relabelling first
  void SettleAfterRelabel();
  void AlignPlacement(QubitRegister &other);
  std::size_t PhysicalIndex(std::size_t data_index) const;

  void AllocateDevice();
  void ReleaseDevice();
  Type *HostAmplitude(std::size_t index) const;
  void BeforeDeviceOp() const;  // write back host-side edits, prefetch managed pages
  void FlushForRead();          // run every queued gate (strict improvement)
  unsigned LocalQubits() const;
  void TimedStart(const std::string &name, std::size_t c, std::size_t t);
  void TimedStop(double algorithmic_bytes, int kind);

  QubitRegister<Type> &operator=(const QubitRegister<Type> &src) { return *this; }
};

template <typename Type>
bool QubitRegister<Type>::do_print_extra_info = false;

template <typename Type>
using BaseType = typename QubitRegister<Type>::BaseType;

}  // namespace iqs

// Derived classes, included at the end of qureg.hpp as the reference does (qureg.hpp:432-433):
// gate/depth counting and automatic insertion of noise gates.
#include "QubitRegisterMetric.hpp"
#include "NoisyQureg.hpp"
