// qureg_permute.cpp -- reordering of the qubits (communication reduction).
//
// Reference behaviour restated: src/qureg_permute.cpp (PermuteQubits :10-41, EmulateSwap :45-52,
// PermuteLocalQubits :55-104, PermuteGlobalQubits :108-187, PermuteByLocalGlobalExchangeOfQubitPairs
// :191-229).  The per-amplitude host loop `state[p2d_new(d2p_old(i))] = old[i]` becomes one bit-
// permutation kernel (csrc/kernels_misc.cu); the rank-to-rank block moves become peer-memory pulls
// over NVLink (csrc/comm.cu).  All of it is pure data movement: bit-exact.
#include "qureg_impl.hpp"

namespace iqs {

using detail::Check;

template <class Type>
void QubitRegister<Type>::PermuteQubits(std::vector<std::size_t> new_map, std::string style_of_map) {
  assert(num_qubits == new_map.size());
  unsigned nprocs = iqs::mpi::Environment::GetStateSize();
  if (nprocs == 1) {
    this->PermuteLocalQubits(new_map, style_of_map);
  } else {
    Permutation &qubit_permutation_old = *qubit_permutation;
    Permutation qubit_permutation_new(new_map, style_of_map);
    std::size_t M = LocalQubits();
    std::vector<std::size_t> int_1_imap, int_2_imap;
    qubit_permutation_old.ObtainIntemediateInverseMaps(qubit_permutation_new.map, M, int_1_imap, int_2_imap);
    // local reshuffle, global reshuffle, then pairwise local<->global exchanges
    this->PermuteLocalQubits(int_1_imap, "inverse");
    this->PermuteGlobalQubits(int_2_imap, "inverse");
    this->PermuteByLocalGlobalExchangeOfQubitPairs(new_map, style_of_map);
  }
}

template <class Type>
void QubitRegister<Type>::EmulateSwap(unsigned qubit_1, unsigned qubit_2) {
  assert(qubit_1 < num_qubits);
  assert(qubit_2 < num_qubits);
  qubit_permutation->ExchangeTwoElements(qubit_1, qubit_2);
}

template <class Type>
void QubitRegister<Type>::PermuteLocalQubits(std::vector<std::size_t> new_map, std::string style_of_map) {
  assert(new_map.size() == this->num_qubits);
  Permutation &old_qubit_permutation = *qubit_permutation;
  Permutation new_qubit_permutation(new_map, style_of_map);
  std::vector<std::size_t> &new_inverse_map = new_qubit_permutation.imap;
  std::vector<std::size_t> &old_inverse_map = qubit_permutation->imap;
  std::size_t M = LocalQubits();
  // the new map must keep every local qubit local and leave the global positions alone
  std::vector<bool> local(new_inverse_map.size(), 0);
  for (unsigned pos = 0; pos < M; ++pos) local[new_inverse_map[pos]] = 1;
  for (unsigned pos = 0; pos < M; ++pos) assert(local[old_inverse_map[pos]] > 0);
  for (unsigned pos = M; pos < num_qubits; ++pos) assert(old_inverse_map[pos] == new_inverse_map[pos]);
  if (old_inverse_map == new_inverse_map) return;

  FlushForRead();
  RestoreCanonicalPlacement();
  BeforeDeviceOp();
  // amplitude i (old data index) moves to program2data_new(data2program_old(i)): bit `pos` of i
  // belongs to qubit old_imap[pos] and lands on position new_map[that qubit]
  std::vector<uint8_t> dst_bit(M);
  for (unsigned pos = 0; pos < M; ++pos) dst_bit[pos] = (uint8_t)new_qubit_permutation.map[old_inverse_map[pos]];
  Check(iqsb_permute_local(dev_, dst_bit.data(), (unsigned)M), "PermuteLocalQubits");
  old_qubit_permutation = new_qubit_permutation;
}

template <class Type>
void QubitRegister<Type>::PermuteGlobalQubits(std::vector<std::size_t> new_map, std::string style_of_map) {
  assert(new_map.size() == this->num_qubits);
  Permutation new_qubit_permutation(new_map, style_of_map);
  std::vector<std::size_t> new_direct_map = new_qubit_permutation.map;
  std::vector<std::size_t> new_inverse_map = new_qubit_permutation.imap;
  std::vector<std::size_t> old_direct_map = qubit_permutation->map;
  std::vector<std::size_t> old_inverse_map = qubit_permutation->imap;
  std::size_t M = LocalQubits();
  std::vector<bool> global(new_inverse_map.size(), 0);
  for (unsigned pos = M; pos < num_qubits; ++pos) global[new_inverse_map[pos]] = 1;
  for (unsigned pos = M; pos < num_qubits; ++pos) assert(global[old_inverse_map[pos]] > 0);
  for (unsigned pos = 0; pos < M; ++pos) assert(old_inverse_map[pos] == new_inverse_map[pos]);
  if (old_inverse_map == new_inverse_map) return;
  assert(iqs::mpi::Environment::GetStateSize() > 1);

  // this rank's shard goes to `destination` and is replaced by `source`'s (permute.cpp:149-166)
  std::size_t myrank = iqs::mpi::Environment::GetStateRank();
  std::size_t source = 0, destination = 0;
  std::size_t glb_start = UL(myrank) * LocalSize();
  for (unsigned pos = M; pos < num_qubits; ++pos) {
    if (check_bit(glb_start, pos) == 1) {
      destination += UL(1) << (new_direct_map[old_inverse_map[pos]] - M);
      source += UL(1) << (old_direct_map[new_inverse_map[pos]] - M);
    }
  }
  FlushForRead();
  RestoreCanonicalPlacement();
  BeforeDeviceOp();
  Check(iqsb_permute_global(dev_, (int)source, (int)destination), "PermuteGlobalQubits");
  qubit_permutation->SetNewPermutationFromMap(new_map, style_of_map);
}

template <class Type>
void QubitRegister<Type>::PermuteByLocalGlobalExchangeOfQubitPairs(std::vector<std::size_t> new_map, std::string style_of_map) {
  Permutation new_qubit_permutation(new_map, style_of_map);
  std::vector<unsigned> exchanged_qubits(num_qubits, 0);
  unsigned num_pairs = 0;
  std::size_t M = LocalQubits();
  for (unsigned qubit = 0; qubit < num_qubits; ++qubit) {
    if (exchanged_qubits[qubit] != 0) continue;
    unsigned old_position = (*qubit_permutation)[qubit];
    unsigned new_position = new_qubit_permutation[qubit];
    if (new_position == old_position) continue;
    // the partner must form a 2-cycle with this qubit, one local and one global
    unsigned partner_qubit = (unsigned)qubit_permutation->Find(new_position);
    assert(exchanged_qubits[partner_qubit] == 0);
    assert(new_qubit_permutation[partner_qubit] == old_position);
    assert((old_position < M) != (new_position < M));
    (void)M;
    ApplySwap(qubit, partner_qubit);  // moves the data
    qubit_permutation->ExchangeTwoElements(qubit, partner_qubit);
    ++num_pairs;
    exchanged_qubits[qubit] = exchanged_qubits[partner_qubit] = num_pairs;
  }
}

template class QubitRegister<ComplexSP>;
template class QubitRegister<ComplexDP>;

}  // namespace iqs
