// qureg_permute.cpp -- reordering of the qubits (communication reduction).
//
// Reference behaviour: src/qureg_permute.cpp (PermuteQubits :10-41, EmulateSwap :45-52,
// PermuteLocalQubits :55-104, PermuteGlobalQubits :108-187, PermuteByLocalGlobalExchangeOfQubitPairs
// :191-229).  There every call moves the data at once: a host loop over all amplitudes into a second
// copy of the state, a Sendrecv loop between ranks, one distributed ApplySwap per local/global pair.
//
// Here a permutation is first of all a RELABELLING.  The register already separates the position a
// program qubit has in the reference's layout (qubit_permutation) from the physical bit that holds
// that position (place_, src/placement.cpp), so PermuteQubits only has to say "position new_map[q] is
// held by whatever bit held q's old position".  The data is brought into the reference's order by the
// one routine that knows how to do that in few passes (RestoreCanonicalPlacement: up to three
// local<->global pairs per NVLink pass, one whole-shard move for the rank bits, one in-place tiled
// bit permutation for the local bits), before the call returns: PermuteQubits is a collective call,
// `psi[i]` afterwards is not -- the reference's own tests read it on the owning rank only, right after
// a permutation (unit_test/include/state_initialization_test.hpp:213-234, found by running that suite on
// 2 GPUs), so nothing collective may be left pending.
// IQS_B200_LAZY_PERMUTE=1 (sharded registers with the placement layer on) leaves the data where it is:
// gates keep running on physical bits, a qubit the program moved to a "local" position is swapped in
// when a gate first needs it, and the order is restored by the first call that reads amplitudes by
// index -- which then has to be made by every rank.  The qubit-reordering example at 35 qubits on
// 8 GPUs runs 1.48 s that way.
// All of it is pure data movement: bit-exact.
#include <cstdlib>

#include "qureg_impl.hpp"

namespace iqs {

using detail::Check;

// Program qubit q is from now on said to live at data position target.map[q]; no amplitude moves.
template <class Type>
void QubitRegister<Type>::Relabel(const Permutation &target) {
  FlushForRead();  // queued gates were resolved against the old map
  const unsigned n = (unsigned)num_qubits;
  std::vector<uint8_t> holder(n);
  std::vector<uint64_t> used(n);
  for (unsigned q = 0; q < n; ++q) {
    const unsigned was = (unsigned)qubit_permutation->map[q], is = (unsigned)target.map[q];
    holder[is] = place_[was];
    used[is] = last_use_[was];
  }
  bool identity = true;
  for (unsigned p = 0; p < n; ++p) {
    place_[p] = holder[p];
    where_[holder[p]] = (uint8_t)p;
    identity = identity && holder[p] == p;
  }
  last_use_ = used;
  moved_ = !identity;
  *qubit_permutation = target;
}

// after a relabelling the data follows now, unless the program opted into the lazy scheme
template <class Type>
void QubitRegister<Type>::SettleAfterRelabel() {
  static const bool lazy = [] {
    const char *e = std::getenv("IQS_B200_LAZY_PERMUTE");
    return e != nullptr && *e != 0 && *e != '0';
  }();
  if (!(placement_ && lazy)) RestoreCanonicalPlacement();
}

template <class Type>
void QubitRegister<Type>::PermuteQubits(std::vector<std::size_t> new_map, std::string style_of_map) {
  assert(num_qubits == new_map.size());
  Permutation target(new_map, style_of_map);
  if (target.map == qubit_permutation->map) return;
  Relabel(target);
  SettleAfterRelabel();
}

template <class Type>
void QubitRegister<Type>::EmulateSwap(unsigned qubit_1, unsigned qubit_2) {
  assert(qubit_1 < num_qubits);
  assert(qubit_2 < num_qubits);
  // the two qubits trade their roles, the amplitudes stay (queued gates hold positions, not qubits)
  qubit_permutation->ExchangeTwoElements(qubit_1, qubit_2);
}

// The three partial permutations of the reference's API.  Each checks what the reference asserts
// about its argument and is then the same relabelling.
template <class Type>
void QubitRegister<Type>::PermuteLocalQubits(std::vector<std::size_t> new_map, std::string style_of_map) {
  assert(new_map.size() == this->num_qubits);
  Permutation target(new_map, style_of_map);
  const std::size_t M = LocalQubits();
  for (std::size_t q = 0; q < num_qubits; ++q) {
    const std::size_t was = qubit_permutation->map[q], is = target.map[q];
    // a local qubit stays local, a global one stays exactly where it is
    assert(was < M ? is < M : is == was);
    (void)was;
    (void)is;
  }
  (void)M;
  if (target.map == qubit_permutation->map) return;
  Relabel(target);
  SettleAfterRelabel();
}

template <class Type>
void QubitRegister<Type>::PermuteGlobalQubits(std::vector<std::size_t> new_map, std::string style_of_map) {
  assert(new_map.size() == this->num_qubits);
  Permutation target(new_map, style_of_map);
  const std::size_t M = LocalQubits();
  for (std::size_t q = 0; q < num_qubits; ++q) {
    const std::size_t was = qubit_permutation->map[q], is = target.map[q];
    // a global qubit stays global, a local one stays exactly where it is
    assert(was >= M ? is >= M : is == was);
    (void)was;
    (void)is;
  }
  (void)M;
  if (target.map == qubit_permutation->map) return;
  assert(iqs::mpi::Environment::GetStateSize() > 1);
  Relabel(target);
  SettleAfterRelabel();
}

template <class Type>
void QubitRegister<Type>::PermuteByLocalGlobalExchangeOfQubitPairs(std::vector<std::size_t> new_map, std::string style_of_map) {
  Permutation target(new_map, style_of_map);
  const std::size_t M = LocalQubits();
  for (std::size_t q = 0; q < num_qubits; ++q) {
    const std::size_t was = qubit_permutation->map[q], is = target.map[q];
    if (was == is) continue;
    // the qubit trades places with exactly one partner, across the local / global border
    const std::size_t partner = qubit_permutation->Find(is);
    assert(target.map[partner] == was);
    assert((was < M) != (is < M));
    (void)partner;
  }
  (void)M;
  if (target.map == qubit_permutation->map) return;
  Relabel(target);  // the pairs travel together, up to three per pass (the reference: one ApplySwap each)
  SettleAfterRelabel();
}

template class QubitRegister<ComplexSP>;
template class QubitRegister<ComplexDP>;

}  // namespace iqs
