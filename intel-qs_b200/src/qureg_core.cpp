// qureg_core.cpp -- construction, HBM allocation, initialisation, the host view of the state and
// the element-wise / reduction utilities of iqs::QubitRegister.
//
// Reference behaviour restated: src/qureg_init.cpp (ctors :22-52, Resize :55-87, Initialize
// :103-157, Allocate :160-190, Initialize(style) :218-360, copy ctor :364-379, toggles :383-447,
// dtor :449-457) and src/qureg_utils.cpp (:17-300).  The loops are CUDA kernels behind the C ABI.
#include <cstdlib>
#include <cstring>

#include "qureg_impl.hpp"

namespace iqs {

using detail::Check;

namespace {
constexpr std::size_t kChunkLog2 = 16;  // host-mirror granularity: 65536 amplitudes = 1 MiB per PCIe round trip (a sweep of operator[] over the state costs 16x fewer of them than with 64 KiB)
bool WantDeviceMemory() {
  const char *e = getenv("IQS_B200_MEM");
  return e && std::string(e) == "device";
}
std::size_t ManagedLimit() {
  const char *e = getenv("IQS_B200_MANAGED_MAX_LOG2");
  int lg = e ? atoi(e) : 31;
  if (lg < 0) lg = 0;
  if (lg > 40) lg = 40;
  return std::size_t(1) << lg;
}
}  // namespace

template <class Type>
unsigned QubitRegister<Type>::LocalQubits() const {
  return (unsigned)(num_qubits - iqs::ilog2(iqs::mpi::Environment::GetStateSize()));
}

// ---------------------------------------------------------------------------------------------
// device allocation / host view
// ---------------------------------------------------------------------------------------------
template <class Type>
void QubitRegister<Type>::AllocateDevice() {
  iqsb_ctx *ctx = iqs::mpi::Environment::Context();
  int nranks = iqs::mpi::Environment::GetStateSize();
  // managed memory gives the host a real pointer; several ranks need IPC-exportable device memory
  // (shards above IQS_B200_MANAGED_MAX_LOG2 amplitudes, default 2^31 = 32 GiB of ComplexDP, always use
  // device memory: a host view of that size is impractical and the driver may refuse the mapping)
  managed_ = (nranks == 1) && !WantDeviceMemory() && !imported_state && LocalSize() <= ManagedLimit();
  std::size_t tmp = (nranks == 1) ? 0 : TmpSize();
  int rc = iqsb_alloc(ctx, LocalSize(), tmp, detail::DType<Type>::value, managed_ ? IQSB_MEM_MANAGED : IQSB_MEM_DEVICE, &dev_);
  if (rc != IQSB_OK && managed_) {  // no managed mapping available: device memory + host mirror
    managed_ = false;
    rc = iqsb_alloc(ctx, LocalSize(), tmp, detail::DType<Type>::value, IQSB_MEM_DEVICE, &dev_);
  }
  Check(rc, "allocating the state vector in HBM");
  if (nranks > 1) Check(iqsb_share(dev_), "publishing the shard to the peer GPUs");
  if (managed_) {
    state = static_cast<Type *>(iqsb_host_ptr(dev_));
  } else if (!imported_state) {
    mirror_ = nullptr;  // allocated on first host access
    state = nullptr;
  }
  host_touched_ = false;
  checked_out_.clear();
  chunk_present_.clear();
  iqs::mpi::detail::LiveRegister live;
  live.self = this;
  live.settle = [](void *self) {
    QubitRegister<Type> *r = static_cast<QubitRegister<Type> *>(self);
    r->FlushForRead();
    r->RestoreCanonicalPlacement();
  };
  live.release = [](void *self) { static_cast<QubitRegister<Type> *>(self)->ReleaseDevice(); };
  iqs::mpi::detail::RegisterLive(live);
}

template <class Type>
void QubitRegister<Type>::ReleaseDevice() {
  iqs::mpi::detail::UnregisterLive(this);
  if (dev_) {
    iqsb_free(dev_);
    dev_ = nullptr;
  }
  if (mirror_ && mirror_owned_) std::free(mirror_);
  mirror_ = nullptr;
  mirror_owned_ = false;
  checked_out_.clear();
  chunk_present_.clear();
}

template <class Type>
Type *QubitRegister<Type>::HostAmplitude(std::size_t index) const {
  assert(index < LocalSize() + TmpSize());
  InvalidateMarginals();  // the caller may write through the pointer
  if (!queue_.empty() || !CanonicalPlacement()) {
    // the host sees the state after every gate issued so far, in the reference's amplitude order
    const_cast<QubitRegister<Type> *>(this)->FlushForRead();
    RestoreCanonicalPlacement();
  }
  if (managed_) {
    if (!host_touched_) {
      Check(iqsb_sync(iqs::mpi::Environment::Context()), "synchronising before a host access");
      host_touched_ = true;
    }
    return state + index;
  }
  if (!mirror_) {
    // virtual reservation: pages are only touched chunk by chunk
    mirror_ = static_cast<Type *>(std::calloc(LocalSize(), sizeof(Type)));
    if (!mirror_) throw std::bad_alloc();
    mirror_owned_ = true;
    const_cast<QubitRegister<Type> *>(this)->state = mirror_;
  }
  if (chunk_present_.empty()) chunk_present_.assign((LocalSize() >> kChunkLog2) + 1, 0);
  std::size_t c = index >> kChunkLog2;
  if (!chunk_present_[c]) {
    std::size_t first = c << kChunkLog2;
    std::size_t count = std::min<std::size_t>(std::size_t(1) << kChunkLog2, LocalSize() - first);
    Check(iqsb_download(dev_, mirror_ + first, first, count), "checking out a chunk of the state");
    chunk_present_[c] = 1;
    checked_out_.push_back(c);
  }
  return mirror_ + index;
}

template <class Type>
void QubitRegister<Type>::BeforeDeviceOp() const {
  InvalidateMarginals();  // whatever follows may change the amplitudes (or already did, from the host)
  if (managed_) {
    if (host_touched_) {
      Check(iqsb_prefetch_device(dev_), "prefetching the state back to HBM");
      host_touched_ = false;
    }
    return;
  }
  if (!checked_out_.empty()) {
    // the host may have written through operator[]: write every checked-out chunk back
    for (std::size_t c : checked_out_) {
      std::size_t first = c << kChunkLog2;
      std::size_t count = std::min<std::size_t>(std::size_t(1) << kChunkLog2, LocalSize() - first);
      Check(iqsb_upload(dev_, mirror_ + first, first, count), "writing back a chunk of the state");
      chunk_present_[c] = 0;
    }
    checked_out_.clear();
  }
}

template <class Type>
Type *QubitRegister<Type>::RawState() {
  FlushForRead();
  RestoreCanonicalPlacement();
  InvalidateMarginals();
  if (managed_) {
    Check(iqsb_sync(iqs::mpi::Environment::Context()), "synchronising before a host access");
    host_touched_ = true;
    raw_exposed_ = true;  // the pointer may be kept and written through at any time (NumPy views)
    return state;
  }
  // device-memory register: hand out the mirror with every chunk checked out
  BeforeDeviceOp();
  for (std::size_t i = 0; i < LocalSize(); i += std::size_t(1) << kChunkLog2) (void)HostAmplitude(i);
  return state;
}

template <class Type>
void QubitRegister<Type>::SyncToHost() {
  FlushForRead();
  RestoreCanonicalPlacement();
  Check(iqsb_sync(iqs::mpi::Environment::Context()), "synchronising");
  if (!managed_ && mirror_) {
    BeforeDeviceOp();
    Check(iqsb_download(dev_, mirror_, 0, LocalSize()), "refreshing the host mirror");
  }
}

template <class Type>
void QubitRegister<Type>::FlushForRead() {
  if (!queue_.empty()) RunQueue(queue_.size());
}

// ---------------------------------------------------------------------------------------------
// constructors
// ---------------------------------------------------------------------------------------------
template <class Type>
QubitRegister<Type>::QubitRegister() {
  // one amplitude equal to 1 (reference qureg_init.cpp:22-52)
  timer = nullptr;
  gate_counter = nullptr;
  qubit_permutation = nullptr;
  imported_state = false;
  specialize = false;
  fusion = false;
  log2llc = 0;
  llc_watermarkbit = 0;
  num_qubits = 1;
  state = nullptr;
  assert(iqs::mpi::Environment::GetStateSize() == 1 && "the default constructor is single-rank only");
  Resize(1UL);
  Check(iqsb_set_amp(dev_, 0, 1., 0.), "initialising the state");
}

template <class Type>
void QubitRegister<Type>::Resize(std::size_t new_num_amplitudes) {
  unsigned log2_nprocs = iqs::ilog2(iqs::mpi::Environment::GetStateSize());
  if (GlobalSize()) assert(GlobalSize() * 2UL == new_num_amplitudes);
  if (dev_) {
    FlushForRead();
    RestoreCanonicalPlacement();
  }
  iqsb_state *old = dev_;
  bool old_managed = managed_;
  std::size_t old_local = local_size_;
  num_qubits = iqs::ilog2(new_num_amplitudes);
  local_size_ = UL(1) << UL(num_qubits - log2_nprocs);
  global_size_ = UL(1) << UL(num_qubits);
  assert(LocalSize() >= 1L);
  if (old) {  // keep the old amplitudes (std::vector::resize semantics of the reference's default build)
    dev_ = nullptr;
    Type *old_mirror = mirror_;
    bool owned = mirror_owned_;
    mirror_ = nullptr;
    mirror_owned_ = false;
    AllocateDevice();
    Check(iqsb_fill_const(dev_, 0., 0.), "clearing the resized state");
    std::vector<Type> keep(old_local);
    Check(iqsb_download(old, keep.data(), 0, old_local), "saving the state before resizing");
    Check(iqsb_upload(dev_, keep.data(), 0, old_local), "restoring the state after resizing");
    iqsb_free(old);
    if (old_mirror && owned) std::free(old_mirror);
    (void)old_managed;
  } else {
    AllocateDevice();
    Check(iqsb_fill_const(dev_, 0., 0.), "clearing the state");
  }
  if (qubit_permutation) delete qubit_permutation;
  qubit_permutation = new Permutation(num_qubits);
  InitPlacement();
}

template <class Type>
void QubitRegister<Type>::AllocateAdditionalQubit() {
  ++num_qubits;
  Resize(UL(1) << UL(num_qubits));
}

template <class Type>
void QubitRegister<Type>::Initialize(std::size_t new_num_qubits, std::size_t tmp_spacesize) {
  unsigned nprocs = iqs::mpi::Environment::GetStateSize();
  unsigned log2_nprocs = iqs::ilog2(nprocs);
  assert(new_num_qubits > log2_nprocs && "Too few qubits for this number of ranks");
  assert(new_num_qubits > 0);
  local_size_ = UL(1) << UL(new_num_qubits - log2_nprocs);
  global_size_ = UL(1) << UL(new_num_qubits);
  assert(LocalSize() > 1);
  std::size_t lcl_size_half = LocalSize() / 2L;
  // same tmp-size policy as the reference (qureg_init.cpp:118-132).  The peer-memory kernels do
  // not need the tmp area; it is only allocated for PermuteGlobalQubits' staging when nranks > 1.
  std::size_t hard_bound = UL(1) << UL(30);
  if (tmp_spacesize == 0 || local_size_ <= tmp_spacesize) this->tmp_spacesize_ = lcl_size_half;
  else if (tmp_spacesize <= hard_bound) {
    assert((lcl_size_half % tmp_spacesize) == 0);
    this->tmp_spacesize_ = tmp_spacesize;
  } else this->tmp_spacesize_ = hard_bound;
  // HBM is the scarce resource: never stage through more than 2^26 amplitudes (1 GiB)
  if (this->tmp_spacesize_ > (UL(1) << 26)) this->tmp_spacesize_ = UL(1) << 26;
  this->num_qubits = new_num_qubits;
  qubit_permutation = new Permutation(new_num_qubits);
  InitPlacement();
  if (do_print_extra_info && !iqs::mpi::Environment::GetStateRank()) printf("Specialization is off\n");
  timer = nullptr;
  gate_counter = nullptr;
}

template <class Type>
void QubitRegister<Type>::Allocate(std::size_t new_num_qubits, std::size_t tmp_spacesize) {
  imported_state = false;
  specialize = false;
  fusion = false;
  log2llc = 0;
  llc_watermarkbit = 0;
  state = nullptr;
  Initialize(new_num_qubits, tmp_spacesize);
  if (do_print_extra_info && !iqs::mpi::Environment::GetStateRank()) {
    double MB = 1024.0 * 1024.0;
    printf("HBM per rank: state = %.2lf MB, staging = %.2lf MB\n", double(LocalSize()) * sizeof(Type) / MB,
           iqs::mpi::Environment::GetStateSize() > 1 ? double(TmpSize()) * sizeof(Type) / MB : 0.0);
  }
  AllocateDevice();
  // IQS_B200_AUTO_FUSION=1: registers start with fusion on, as if TurnOnFusion() had been called --
  // unchanged programs get one HBM sweep per run of gates; results are bit-identical, reads flush.
  if (const char *e = getenv("IQS_B200_AUTO_FUSION"))
    if (e[0] == '1' && LocalQubits() >= 2) {
      log2llc = LocalQubits();
      fusion = true;
    }
}

template <class Type>
QubitRegister<Type>::QubitRegister(std::size_t new_num_qubits, Type *state_, std::size_t tmp_spacesize) {
  imported_state = true;
  specialize = false;
  fusion = false;
  log2llc = 0;
  llc_watermarkbit = 0;
  Initialize(new_num_qubits, tmp_spacesize);
  AllocateDevice();
  mirror_ = state_;
  mirror_owned_ = false;
  this->state = state_;
  Check(iqsb_upload(dev_, state_, 0, LocalSize()), "uploading the imported state");
}

template <class Type>
QubitRegister<Type>::QubitRegister(std::size_t new_num_qubits, std::string style, std::size_t base_index, std::size_t tmp_spacesize) {
  Allocate(new_num_qubits, tmp_spacesize);
  Initialize(style, base_index);
}

template <class Type>
void QubitRegister<Type>::Initialize(std::string style, std::size_t base_index) {
  BeforeDeviceOp();
  queue_.clear();  // the state is overwritten: pending gates and qubit moves are moot
  for (std::size_t p = 0; p < place_.size(); ++p) place_[p] = where_[p] = (uint8_t)p;
  moved_ = false;
  Check(iqsb_fill_const(dev_, 0., 0.), "clearing the state");
  if (style == "rand") {
    // Same stream layout as the reference (qureg_init.cpp:256-332): numbers are drawn on the host
    // from the pool stream (base_index == 0, skipping 2*rank*L) or the local stream, uploaded, then
    // the state is normalised on the device.
    assert(rng_ptr_ != nullptr);
    assert(rng_ptr_->GetSeed() != 0);
    assert(base_index == 0 || base_index == (std::size_t)iqs::mpi::Environment::GetNumStates());
    std::size_t myrank = iqs::mpi::Environment::GetStateRank();
    iqs::RandomNumberGenerator<BaseType> draw(rng_ptr_);
    const char *stream = base_index > 0 ? "local" : "pool";
    if (base_index == 0) draw.SkipAhead(2UL * myrank * LocalSize(), "pool");
    const std::size_t block = std::size_t(1) << 20;
    std::vector<BaseType> buf(2 * std::min(block, LocalSize()));
    for (std::size_t first = 0; first < LocalSize(); first += block) {
      std::size_t count = std::min(block, LocalSize() - first);
      draw.UniformRandomNumbers(buf.data(), 2UL * count, -1., 1., stream);
      Check(iqsb_upload(dev_, buf.data(), first, count), "uploading random amplitudes");
    }
    if (base_index > 0) rng_ptr_->SkipAhead(2UL * LocalSize(), "local");
    else rng_ptr_->SkipAhead(2UL * GlobalSize(), "pool");
    this->Normalize();
  } else if (style == "base") {
    assert(base_index < GlobalSize());
    this->SetGlobalAmplitude(base_index, Type(1.0, 0.0));
  } else if (style == "++++") {
    Type amplitude = {BaseType(1. / std::sqrt(GlobalSize())), 0.};
    this->InitializationWithSameAmplitudeEverywhere(amplitude);
  }
  iqs::mpi::StateBarrier();
}

template <class Type>
QubitRegister<Type>::QubitRegister(const QubitRegister &in) {
  Allocate(in.num_qubits, in.TmpSize());
  const_cast<QubitRegister &>(in).FlushForRead();
  in.BeforeDeviceOp();
  Check(iqsb_copy(dev_, in.dev_), "copying a register");
  *qubit_permutation = *(in.qubit_permutation);
  place_ = in.place_;  // the shard is copied as it lies, with the source's placement
  where_ = in.where_;
  moved_ = in.moved_;
}

template <class Type>
QubitRegister<Type>::~QubitRegister() {
  try {
    if (imported_state && dev_ && mirror_) {
      FlushForRead();
      RestoreCanonicalPlacement();
      BeforeDeviceOp();
      iqsb_download(dev_, mirror_, 0, LocalSize());
    }
  } catch (...) {
  }
  ReleaseDevice();
  if (timer != nullptr) delete timer;
  if (gate_counter != nullptr) delete gate_counter;
  if (qubit_permutation != nullptr) delete qubit_permutation;
}

template <class Type>
void QubitRegister<Type>::TurnOnSpecialize() {
  if (do_print_extra_info && !iqs::mpi::Environment::GetStateRank()) printf("Specialization is on\n");
  specialize = true;
}
template <class Type>
void QubitRegister<Type>::TurnOffSpecialize() {
  if (do_print_extra_info && !iqs::mpi::Environment::GetStateRank()) printf("Specialization is off\n");
  specialize = false;
}
template <class Type>
void QubitRegister<Type>::TurnOnSpecializeV2() {
  if (do_print_extra_info && !iqs::mpi::Environment::GetStateRank()) printf("Specialization v2 is on\n");
  specialize2 = true;
}
template <class Type>
void QubitRegister<Type>::TurnOffSpecializeV2() {
  if (do_print_extra_info && !iqs::mpi::Environment::GetStateRank()) printf("Specialization v2 is off\n");
  specialize2 = false;
}

// ---------------------------------------------------------------------------------------------
// utilities (reference src/qureg_utils.cpp)
// ---------------------------------------------------------------------------------------------
template <class Type>
bool QubitRegister<Type>::operator==(const QubitRegister &rhs) {
  assert(rhs.GlobalSize() == GlobalSize());
  assert(rhs.qubit_permutation->map == qubit_permutation->map);
  FlushForRead();
  const_cast<QubitRegister &>(rhs).FlushForRead();
  AlignPlacement(const_cast<QubitRegister &>(rhs));
  BeforeDeviceOp();
  rhs.BeforeDeviceOp();
  int eq = 0;
  Check(iqsb_equal(dev_, rhs.dev_, &eq), "comparing registers");
  return eq != 0;  // local comparison, as in the reference (:17-31)
}

template <class Type>
typename QubitRegister<Type>::BaseType QubitRegister<Type>::MaxAbsDiff(QubitRegister &x, Type sfactor) {
  assert(LocalSize() == x.LocalSize());
  assert(x.qubit_permutation->map == qubit_permutation->map);
  FlushForRead();
  x.FlushForRead();
  AlignPlacement(x);
  BeforeDeviceOp();
  x.BeforeDeviceOp();
  double s[2] = {sfactor.real(), sfactor.imag()}, v = 0;
  Check(iqsb_maxabsdiff(dev_, x.dev_, s, &v), "MaxAbsDiff");
  iqs::mpi::AllreduceDouble(&v, 1, iqs::mpi::MAX);
  return (BaseType)v;
}

template <class Type>
typename QubitRegister<Type>::BaseType QubitRegister<Type>::MaxL2NormDiff(QubitRegister &x) {
  assert(LocalSize() == x.LocalSize());
  assert(x.qubit_permutation->map == qubit_permutation->map);
  FlushForRead();
  x.FlushForRead();
  AlignPlacement(x);
  BeforeDeviceOp();
  x.BeforeDeviceOp();
  double v = 0;
  Check(iqsb_l2diff(dev_, x.dev_, &v), "MaxL2NormDiff");
  iqs::mpi::AllreduceDouble(&v, 1, iqs::mpi::MAX);  // the reference reduces with MAX (:151)
  return (BaseType)v;
}

template <class Type>
Type QubitRegister<Type>::GetGlobalAmplitude(std::size_t global_index) const {
  assert(global_index < global_size_);
  const_cast<QubitRegister *>(this)->FlushForRead();
  BeforeDeviceOp();
  global_index = PhysicalIndex(qubit_permutation->program2data_(global_index));
  std::size_t hosting_rank = global_index / local_size_, local_index = global_index % local_size_;
  double v[2] = {0, 0};
  if ((int)hosting_rank == iqs::mpi::Environment::GetStateRank()) Check(iqsb_get_amp(dev_, local_index, &v[0], &v[1]), "GetGlobalAmplitude");
  iqs::mpi::BcastDouble(v, 2, (int)hosting_rank);
  return Type((BaseType)v[0], (BaseType)v[1]);
}

template <class Type>
void QubitRegister<Type>::SetGlobalAmplitude(std::size_t global_index, Type value) {
  assert(global_index < global_size_);
  FlushForRead();
  BeforeDeviceOp();
  global_index = PhysicalIndex(qubit_permutation->program2data_(global_index));
  std::size_t hosting_rank = global_index / local_size_, local_index = global_index % local_size_;
  if ((int)hosting_rank == iqs::mpi::Environment::GetStateRank())
    Check(iqsb_set_amp(dev_, local_index, value.real(), value.imag()), "SetGlobalAmplitude");
}

template <class Type>
void QubitRegister<Type>::Normalize() {
  BaseType global_norm = ComputeNorm();
  Type inverse_global_norm(1 / global_norm, 0);
  AmplitudeWiseScalarMultiplication(inverse_global_norm);
}

template <class Type>
void QubitRegister<Type>::InitializationWithSameAmplitudeEverywhere(Type amplitude) {
  BeforeDeviceOp();
  queue_.clear();  // every amplitude is overwritten (and the constant state has no qubit order)
  Check(iqsb_fill_const(dev_, amplitude.real(), amplitude.imag()), "InitializationWithSameAmplitudeEverywhere");
}

template <class Type>
void QubitRegister<Type>::AmplitudeWiseScalarMultiplication(Type factor) {
  FlushForRead();
  BeforeDeviceOp();
  // state[i] = state[i] * factor for every i, also when factor == 1 (qureg_utils.cpp:188-196):
  // multiplying by (1,0) changes no value, so the engine's skip in iqsb_scale is value-identical.
  double f[2] = {factor.real(), factor.imag()};
  Check(iqsb_scale(dev_, f, 0, LocalSize()), "AmplitudeWiseScalarMultiplication");
}

template <class Type>
void QubitRegister<Type>::AmplitudeWiseSum(QubitRegister<Type> &psi, Type factor) {
  assert(LocalSize() == psi.LocalSize());
  FlushForRead();
  psi.FlushForRead();
  AlignPlacement(psi);
  BeforeDeviceOp();
  psi.BeforeDeviceOp();
  double f[2] = {factor.real(), factor.imag()};
  Check(iqsb_axpy(dev_, psi.dev_, f), "AmplitudeWiseSum");
}

template <class Type>
typename QubitRegister<Type>::BaseType QubitRegister<Type>::ComputeNorm() {
  FlushForRead();
  BeforeDeviceOp();
  double v = 0;
  Check(iqsb_norm2(dev_, &v), "ComputeNorm");
  iqs::mpi::AllreduceDouble(&v, 1, iqs::mpi::SUM);
  return (BaseType)std::sqrt(v);
}

template <class Type>
Type QubitRegister<Type>::ComputeOverlap(QubitRegister<Type> &psi) {
  assert(LocalSize() == psi.LocalSize());
  assert(psi.qubit_permutation->map == qubit_permutation->map);
  FlushForRead();
  psi.FlushForRead();
  AlignPlacement(psi);
  BeforeDeviceOp();
  psi.BeforeDeviceOp();
  double v[2] = {0, 0};
  Check(iqsb_overlap(dev_, psi.dev_, v), "ComputeOverlap");
  iqs::mpi::AllreduceDouble(v, 2, iqs::mpi::SUM);
  return Type((BaseType)v[0], (BaseType)v[1]);
}

template <class Type>
double QubitRegister<Type>::Entropy() {
  FlushForRead();
  BeforeDeviceOp();
  double s[11];
  Check(iqsb_entropy_stats(dev_, s), "Entropy");
  iqs::mpi::AllreduceDouble(s, 1, iqs::mpi::SUM);
  return s[0] / std::log(2.0);
}

template <class Type>
std::vector<double> QubitRegister<Type>::GoogleStats() {
  FlushForRead();
  BeforeDeviceOp();
  double s[11];
  Check(iqsb_entropy_stats(dev_, s), "GoogleStats");
  double two2n = double(GlobalSize());
  std::vector<double> stats;
  // local moments are scaled before the reduction, as in the reference (qureg_utils.cpp:418-440)
  double factorial = 1.0;
  for (int i = 0; i < 9; ++i) {
    int k = i + 2;
    factorial *= double(k);
    s[2 + i] *= std::pow(two2n, double(k - 1)) / factorial;
  }
  iqs::mpi::AllreduceDouble(s, 11, iqs::mpi::SUM);
  stats.push_back(s[0] / std::log(2.0));
  stats.push_back(s[1] / std::log(2.0) / two2n);
  for (int i = 0; i < 9; ++i) stats.push_back(s[2 + i]);
  return stats;
}

// ---------------------------------------------------------------------------------------------
// statistics
// ---------------------------------------------------------------------------------------------
template <class Type>
void QubitRegister<Type>::EnableStatistics() {
  int myrank = iqs::mpi::Environment::GetStateRank(), nprocs = iqs::mpi::Environment::GetStateSize();
  assert(timer == nullptr);
  timer = new Timer(num_qubits, myrank, nprocs);
  assert(gate_counter == nullptr);
  gate_counter = new GateCounter(num_qubits);
}
template <class Type>
void QubitRegister<Type>::GetStatistics() {
  assert(timer);
  timer->Breakdown();
  assert(gate_counter);
  gate_counter->Breakdown();
}
template <class Type>
void QubitRegister<Type>::DisableStatistics() {
  assert(timer);
  delete timer;
  timer = nullptr;
  assert(gate_counter);
  delete gate_counter;
  gate_counter = nullptr;
}
template <class Type>
void QubitRegister<Type>::ResetStatistics() {
  assert(timer);
  timer->Reset();
  assert(gate_counter);
  gate_counter->Reset();
}

template <class Type>
void QubitRegister<Type>::TimedStart(const std::string &name, std::size_t c, std::size_t t) {
  if (!timer) return;
  timer->Start(name, c, t);
  Check(iqsb_timer_start(iqs::mpi::Environment::Context()), "timer");
}
// kind: 0 = sn (two-pointer / scale), 1 = dn (1-qubit), 2 = tn (controlled / swap), 3 = cm (NVLink)
template <class Type>
void QubitRegister<Type>::TimedStop(double bytes, int kind) {
  if (!timer) return;
  double ms = 0;
  Check(iqsb_timer_stop(iqs::mpi::Environment::Context(), &ms), "timer");
  double s = ms * 1e-3, bw = s > 0 ? bytes / s : 0;
  if (kind == 0) timer->record_sn(s, bw);
  else if (kind == 1) timer->record_dn(s, bw);
  else if (kind == 2) timer->record_tn(s, bw);
  else timer->record_cm(s, bw);
  timer->Stop();
}

template class QubitRegister<ComplexSP>;
template class QubitRegister<ComplexDP>;

}  // namespace iqs
