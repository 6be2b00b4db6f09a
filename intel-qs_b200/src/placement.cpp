// placement.cpp -- the gate queue and the placement layer of iqs::QubitRegister.
//
// The reference leaves a qubit where qubit_permutation puts it: with P ranks the top log2(P) positions
// are rank bits, and every non-diagonal gate on one of them pays a pairwise exchange of half a shard
// each way (HP_Distrpair, reference src/qureg_apply1qubitgate.cpp:115-160,
// src/qureg_applyctrl1qubitgate.cpp:24-220).  Its remedy, PermuteQubits / EmulateSwap
// (src/qureg_permute.cpp:10-52), is left to the user.  Here the engine does it behind the API:
//
//   * `place_[position]` = the physical bit of the distributed index that holds the position now
//     (identity = the reference's layout).  Kernels are issued on physical bits.
//   * gates wait in `queue_` (fusion on: until a flush; fusion off and several ranks: for a look-ahead
//     window), so when a gate needs a position that sits in a rank bit the planner
//     (iqsb_plan_placement, csrc/placement.cu) sees what comes next and trades it -- together with the
//     other rank-bit positions needed soon -- for the local positions whose next use lies farthest
//     ahead, in ONE multi-bit exchange over NVLink (iqsb_exchange_bits, csrc/exchange.cu).  The
//     qubit then STAYS local.
//   * anything that exposes the raw amplitude order (operator[], RawState, dumps, PermuteQubits,
//     two registers with different placements) first restores the identity placement; reductions and
//     Get/SetGlobalAmplitude translate indices instead.
//
// Data movement is exact, so results are the ones the reference's layout gives, bit for bit.
// IQS_B200_PLACEMENT=0 switches the layer off (gates on rank bits then run as the peer-memory pair
// kernels of csrc/comm.cu); IQS_B200_LOOKAHEAD sets the look-ahead window (default 128 gates).
#include <cstdlib>
#include <cstring>

#include "qureg_impl.hpp"

namespace iqs {

using detail::Check;

namespace {
constexpr std::size_t kPlanWindow = 512;  // gates shown to the planner
constexpr std::size_t kMaxQueued = 4000;  // fused window bound

unsigned EnvUnsigned(const char *name, unsigned fallback) {
  const char *e = getenv(name);
  if (!e || !*e) return fallback;
  long v = atol(e);
  return v < 0 ? 0u : (unsigned)v;
}
}  // namespace

template <class Type>
void QubitRegister<Type>::InitPlacement() {
  const unsigned n = (unsigned)num_qubits;
  place_.resize(n);
  where_.resize(n);
  for (unsigned p = 0; p < n; ++p) place_[p] = where_[p] = (uint8_t)p;
  last_use_.assign(n, 0);
  use_clock_ = 0;
  moved_ = false;
  queue_.clear();
  const int nranks = iqs::mpi::Environment::GetStateSize();
  placement_ = nranks > 1 && LocalQubits() >= 2 && EnvUnsigned("IQS_B200_PLACEMENT", 1) != 0;
  lookahead_ = placement_ ? EnvUnsigned("IQS_B200_LOOKAHEAD", 128) : 0;
}

template <class Type>
void QubitRegister<Type>::SwapPlacement(unsigned a, unsigned b) {
  std::swap(place_[a], place_[b]);
  where_[place_[a]] = (uint8_t)a;
  where_[place_[b]] = (uint8_t)b;
  moved_ = true;
}

// bit P of a data index (position space) -> physical bit place_[P]
template <class Type>
std::size_t QubitRegister<Type>::PhysicalIndex(std::size_t data_index) const {
  if (CanonicalPlacement()) return data_index;
  std::size_t out = 0;
  for (std::size_t p = 0; p < place_.size(); ++p)
    if ((data_index >> p) & 1) out |= std::size_t(1) << place_[p];
  return out;
}

// ---------------------------------------------------------------------------------------------
// queue
// ---------------------------------------------------------------------------------------------
template <class Type>
void QubitRegister<Type>::Enqueue(int kind, unsigned control_position, unsigned target_position, TM2x2<Type> const &m) {
  QueuedGate g;
  g.kind = kind;
  g.control = control_position;
  g.target = target_position;
  detail::M8(m, g.m);
  g.diagonal = g.m[2] == 0. && g.m[3] == 0. && g.m[4] == 0. && g.m[5] == 0.;
  queue_.push_back(g);
  ++use_clock_;
  last_use_[target_position] = use_clock_;
  if (kind == 1) last_use_[control_position] = use_clock_;
  if (fusion) {
    if (queue_.size() >= kMaxQueued) RunQueue(queue_.size());
  } else if (queue_.size() >= 2 * (std::size_t)lookahead_) {
    RunQueue(queue_.size() - lookahead_);  // the oldest gates run; the planner still sees `lookahead_` gates ahead
  }
}

// Run the first `count` queued gates in order.  A gate whose target sits in a rank bit (and is not
// diagonal) first gets the target swapped in; what lies between such gates runs fused (one HBM sweep
// per run of gates) when fusion is on, gate by gate otherwise.
template <class Type>
void QubitRegister<Type>::RunQueue(std::size_t count) {
  if (count > queue_.size()) count = queue_.size();
  if (count == 0) return;
  BeforeDeviceOp();
  const unsigned M = LocalQubits();
  std::size_t i = 0;
  while (i < count) {
    std::size_t j = i;
    while (j < count && !(Phys(queue_[j].target) >= M && !queue_[j].diagonal)) ++j;
    if (j > i) {
      if (fusion) ExecFusedRange(i, j);
      else
        for (std::size_t k = i; k < j; ++k) {
          const QueuedGate &g = queue_[k];
          if (g.kind == 0) ExecGate1(g.target, g.m, g.diagonal, 0UL, LocalSize(), std::string());
          else ExecCGate1(g.control, g.target, g.m, g.diagonal, 0UL, LocalSize(), std::string(), nullptr);
        }
    }
    if (j == count) break;
    if (placement_) {
      BringLocal(j, uint64_t(1) << queue_[j].target);
      i = j;  // the gate no longer blocks
    } else {
      // placement layer off: the gate runs as a peer-memory pair kernel over NVLink (csrc/comm.cu)
      const QueuedGate &g = queue_[j];
      if (g.kind == 0) ExecGate1(g.target, g.m, g.diagonal, 0UL, LocalSize(), std::string());
      else ExecCGate1(g.control, g.target, g.m, g.diagonal, 0UL, LocalSize(), std::string(), nullptr);
      i = j + 1;
    }
  }
  queue_.erase(queue_.begin(), queue_.begin() + (std::ptrdiff_t)count);
}

// queue_[first, last): every non-diagonal target is local.  Gates touching rank bits turn into
// per-rank constants: a global control selects the ranks on which the gate exists at all, a diagonal
// gate on a global target multiplies by the matrix entry of this rank's bit.
template <class Type>
void QubitRegister<Type>::ExecFusedRange(std::size_t first, std::size_t last) {
  const unsigned M = LocalQubits();
  const unsigned myrank = (unsigned)iqs::mpi::Environment::GetStateRank();
  auto rank_bit = [&](unsigned physical) { return (myrank >> (physical - M)) & 1u; };
  std::vector<iqsb_fgate> batch;
  batch.reserve(last - first);
  for (std::size_t k = first; k < last; ++k) {
    const QueuedGate &q = queue_[k];
    const unsigned T = Phys(q.target);
    iqsb_fgate g;
    memset(&g, 0, sizeof(g));
    memcpy(g.m, q.m, sizeof(g.m));
    if (q.kind == 0) {
      if (T < M) {
        g.kind = 0;
        g.target = (int)T;
      } else {  // diagonal on a rank bit: every amplitude times m00 or m11
        const double *s = rank_bit(T) ? &q.m[6] : &q.m[0];
        if (detail::IsOne(s[0], s[1])) continue;
        g.kind = 0;
        g.target = 0;
        g.m[0] = g.m[6] = s[0];
        g.m[1] = g.m[7] = s[1];
        g.m[2] = g.m[3] = g.m[4] = g.m[5] = 0.;
      }
    } else {
      const unsigned C = Phys(q.control);
      if (C >= M && !rank_bit(C)) continue;  // the gate does not exist on this rank
      if (T < M) {
        g.target = (int)T;
        if (C >= M) g.kind = 0;
        else { g.kind = 1; g.control = (int)C; }
      } else {  // diagonal, target on a rank bit: diag(1, s) on the control (or a plain factor)
        const double *s = rank_bit(T) ? &q.m[6] : &q.m[0];
        if (detail::IsOne(s[0], s[1])) continue;
        g.kind = 0;
        g.m[2] = g.m[3] = g.m[4] = g.m[5] = 0.;
        g.m[6] = s[0];
        g.m[7] = s[1];
        if (C >= M) { g.target = 0; g.m[0] = s[0]; g.m[1] = s[1]; }
        else { g.target = (int)C; g.m[0] = 1.; g.m[1] = 0.; }
      }
    }
    batch.push_back(g);
  }
  if (batch.empty()) return;
  if (last - first == 1) {  // a window of one gate is applied as a plain gate (reference fusion.cpp:75)
    const QueuedGate &q = queue_[first];
    if (q.kind == 0) ExecGate1(q.target, q.m, q.diagonal, 0UL, LocalSize(), std::string());
    else ExecCGate1(q.control, q.target, q.m, q.diagonal, 0UL, LocalSize(), std::string(), nullptr);
    return;
  }
  TimedStart("FUSED(" + iqs::toString(batch.size()) + ")", 0, 999999);
  const std::size_t kMaxPerCall = 4096;
  for (std::size_t b = 0; b < batch.size(); b += kMaxPerCall) {
    int cnt = (int)std::min(kMaxPerCall, batch.size() - b);
    Check(iqsb_fused(dev_, batch.data() + b, cnt), "fused gate batch");
  }
  TimedStop(2.0 * sizeof(Type) * double(LocalSize()), 1);
}

// ---------------------------------------------------------------------------------------------
// placement changes
// ---------------------------------------------------------------------------------------------
// Make every position of protect_mask local (and keep it so), using the queued gates from
// `queue_from` on as the planner's view of the future.
template <class Type>
void QubitRegister<Type>::BringLocal(std::size_t queue_from, uint64_t protect_mask) {
  const unsigned n = (unsigned)num_qubits, M = LocalQubits();
  auto pending = [&]() {
    for (unsigned p = 0; p < n; ++p)
      if (((protect_mask >> p) & 1) && place_[p] >= M) return true;
    return false;
  };
  if (!pending()) return;
  std::vector<iqsb_pgate> view;
  const std::size_t end = std::min(queue_.size(), queue_from + kPlanWindow);
  view.reserve(end > queue_from ? end - queue_from : 0);
  for (std::size_t k = queue_from; k < end; ++k) {
    iqsb_pgate g;
    g.kind = queue_[k].kind;
    g.control = (int)queue_[k].control;
    g.target = (int)queue_[k].target;
    g.diagonal = queue_[k].diagonal ? 1 : 0;
    view.push_back(g);
  }
  // low local bits are kept out of the exchange (short runs on the link) unless the shard is tiny
  unsigned min_evict = M >= 12 ? 5u : (M >= 7 ? 2u : 0u);
  for (int round = 0; pending(); ++round) {
    unsigned evict[3], bring[3];
    int k = 0;
    Check(iqsb_plan_placement(place_.data(), n, M, view.data(), (int)view.size(), protect_mask, last_use_.data(), min_evict, evict, bring, &k),
          "planning a qubit exchange");
    if (k == 0) {
      if (min_evict == 0 || round > 8) throw std::runtime_error("iqs (B200 engine): no local qubit can make room for a global one");
      min_evict = 0;
      continue;
    }
    unsigned lpos[3], gpos[3];
    for (int j = 0; j < k; ++j) {
      lpos[j] = place_[evict[j]];
      gpos[j] = place_[bring[j]];
    }
    TimedStart("XCHG(" + iqs::toString(k) + ")", gpos[0], lpos[0]);
    Check(iqsb_exchange_bits(dev_, M, k, lpos, gpos), "exchanging local and global qubits");
    TimedStop((1.0 - 1.0 / double(1 << k)) * sizeof(Type) * double(LocalSize()), 3);
    for (int j = 0; j < k; ++j) SwapPlacement(evict[j], bring[j]);
    ++exchanges_;
    exchanged_bits_ += (uint64_t)k;
  }
}

// Undo every move of the placement layer: afterwards position P is held by physical bit P, i.e. the
// shards are laid out exactly as the reference lays them out (src/qureg_init.cpp:97-107).
template <class Type>
void QubitRegister<Type>::RestoreCanonicalPlacement() const {
  if (CanonicalPlacement()) return;
  QubitRegister<Type> *self = const_cast<QubitRegister<Type> *>(this);
  const unsigned n = (unsigned)num_qubits, M = LocalQubits();
  BeforeDeviceOp();
  // 1. rank bits: position g belongs in rank bit g.  While some such position sits in a local bit,
  //    trade it for whatever occupies its rank bit (several at once).
  for (int round = 0; round < 64; ++round) {
    unsigned lpos[3], gpos[3], pa[3], pb[3];
    int k = 0;
    for (unsigned g = M; g < n && k < 3 && k < (int)M - 1; ++g)
      if (where_[g] != g && place_[g] < M) {
        lpos[k] = place_[g];
        gpos[k] = g;
        pa[k] = g;
        pb[k] = where_[g];
        ++k;
      }
    if (k == 0) break;
    Check(iqsb_exchange_bits(dev_, M, k, lpos, gpos), "restoring the qubit placement");
    for (int j = 0; j < k; ++j) self->SwapPlacement(pa[j], pb[j]);
    ++exchanges_;
    exchanged_bits_ += (uint64_t)k;
  }
  // (the exchange splits its work between the two partners along one more local bit: a register with a
  //  single local qubit -- two amplitudes per GPU -- cannot trade it)
  for (unsigned g = M; g < n; ++g)
    if (place_[g] < M) throw std::runtime_error("iqs: moving a qubit across the local/global border needs at least two local qubits per rank");
  // 2. rank bits permuted among themselves: one whole-shard move (the rank permutation of
  //    PermuteGlobalQubits, reference src/qureg_permute.cpp:149-185)
  bool ranks_ok = true;
  for (unsigned g = M; g < n; ++g) ranks_ok = ranks_ok && where_[g] == g;
  if (!ranks_ok) {
    // the content of rank bit g is position where_[g], which belongs in rank bit where_[g]; every
    // rank derives its source, destination and the path from this table (no collective to agree)
    std::vector<uint8_t> dst_rank_bit(n - M);
    for (unsigned g = M; g < n; ++g) dst_rank_bit[g - M] = (uint8_t)(where_[g] - M);
    Check(iqsb_permute_global_bits(dev_, dst_rank_bit.data(), n - M), "restoring the order of the global qubits");
    for (unsigned g = M; g < n; ++g) place_[g] = where_[g] = (uint8_t)g;
  }
  // 3. local bits: one bit permutation of the local index (in-place tile phases)
  std::vector<uint8_t> dst_bit(M);
  bool identity = true;
  for (unsigned b = 0; b < M; ++b) {
    dst_bit[b] = where_[b];  // the content of bit b is position where_[b], whose home is bit where_[b]
    identity = identity && dst_bit[b] == b;
  }
  if (!identity) Check(iqsb_permute_local(dev_, dst_bit.data(), M), "restoring the order of the local qubits");
  for (unsigned p = 0; p < n; ++p) place_[p] = where_[p] = (uint8_t)p;
  moved_ = false;
}

// Two registers meet in one kernel (overlap, differences, axpy, ==): same placement needed.
template <class Type>
void QubitRegister<Type>::AlignPlacement(QubitRegister &other) {
  if (place_ == other.place_) return;
  RestoreCanonicalPlacement();
  other.RestoreCanonicalPlacement();
}

template class QubitRegister<ComplexSP>;
template class QubitRegister<ComplexDP>;

}  // namespace iqs
