// placement.cu -- which qubits live in the rank bits: the planning half of the placement layer
// ("global qubits as a cache").  Pure host code, no device access: it is exercised on CPU by
// tests/test_placement_plan.py and called by the host library (intel-qs_b200/src/placement.cpp)
// right before iqsb_exchange_bits.
//
// Reference context: the reference keeps the data where qubit_permutation says and pays one
// HP_Distrpair exchange (16*L bytes each way) for EVERY non-diagonal gate on a global qubit
// (src/qureg_apply1qubitgate.cpp:115-160); its own remedy is the user-driven PermuteQubits /
// EmulateSwap (src/qureg_permute.cpp:10-52, examples/communication_reduction_via_qubit_reordering.cpp).
// Here the engine does that reordering itself, behind the API: a position (data qubit) is held by a
// PHYSICAL bit -- a bit of the local index or a rank bit -- and when an upcoming gate needs a position
// that sits in a rank bit, the planner picks local positions to trade for it (and for the other
// rank-bit positions needed soon) so that ONE multi-bit exchange serves a whole stretch of the
// circuit.  The choice is Belady's: evict the positions whose next "hard" use lies farthest ahead.
//
// "Hard" use of a position: target of a non-diagonal gate (needs both halves of the pair on one
// GPU).  Controls and diagonal gates are free on a rank bit (the gate turns into a per-rank
// constant or a per-rank no-op).
#include <string.h>

#include "iqsb_internal.cuh"

extern "C" int iqsb_plan_placement(const uint8_t *place, unsigned n, unsigned M, const iqsb_pgate *gates, int ngates, uint64_t protect_mask,
                                   const uint64_t *last_use, unsigned min_evict_bit, unsigned *evict_pos, unsigned *bring_pos, int *k_out) {
  IQSB_REQUIRE(place && evict_pos && bring_pos && k_out && (gates || ngates == 0), "iqsb_plan_placement: null argument");
  IQSB_REQUIRE(n <= 62 && M < n, "iqsb_plan_placement: nothing is global (n = %u, M = %u)", n, M);
  const int kInf = 0x7fffffff;
  int first_hard[64], first_any[64];
  for (unsigned p = 0; p < n; ++p) first_hard[p] = first_any[p] = kInf;
  for (int i = 0; i < ngates; ++i) {
    const iqsb_pgate &g = gates[i];
    IQSB_REQUIRE(g.target >= 0 && (unsigned)g.target < n && (g.kind == 0 || (g.control >= 0 && (unsigned)g.control < n)), "iqsb_plan_placement: bad gate %d", i);
    if (!g.diagonal && first_hard[g.target] == kInf) first_hard[g.target] = i;
    if (first_any[g.target] == kInf) first_any[g.target] = i;
    if (g.kind == 1 && first_any[g.control] == kInf) first_any[g.control] = i;
  }
  // positions in rank bits with a hard use ahead, soonest first; protected positions count as needed now
  unsigned need[64];
  int nneed = 0;
  for (unsigned p = 0; p < n; ++p)
    if (place[p] >= M) {
      if ((protect_mask >> p) & 1) first_hard[p] = -1;
      if (first_hard[p] != kInf) need[nneed++] = p;
    }
  for (int i = 0; i < nneed; ++i)
    for (int j = i + 1; j < nneed; ++j)
      if (first_hard[need[j]] < first_hard[need[i]]) { unsigned t = need[i]; need[i] = need[j]; need[j] = t; }
  // eviction candidates: local positions on physical bits >= min_evict_bit, farthest next hard use first;
  // ties: farthest next use of any kind, then least recently used, then the highest physical bit
  unsigned cand[64];
  int ncand = 0;
  for (unsigned p = 0; p < n; ++p)
    if (place[p] < M && place[p] >= min_evict_bit && !((protect_mask >> p) & 1)) cand[ncand++] = p;
  auto better = [&](unsigned a, unsigned b) {  // a is a better victim than b
    if (first_hard[a] != first_hard[b]) return first_hard[a] > first_hard[b];
    if (first_any[a] != first_any[b]) return first_any[a] > first_any[b];
    if (last_use && last_use[a] != last_use[b]) return last_use[a] < last_use[b];
    return place[a] > place[b];
  };
  for (int i = 0; i < ncand; ++i)
    for (int j = i + 1; j < ncand; ++j)
      if (better(cand[j], cand[i])) { unsigned t = cand[i]; cand[i] = cand[j]; cand[j] = t; }
  int kmax = 3;
  if ((int)M - 1 < kmax) kmax = (int)M - 1;
  int k = 0;
  for (int i = 0; i < nneed && i < ncand && k < kmax; ++i) {
    // a trade pays only if the incoming position is needed before the evicted one (the first trade is forced)
    if (i > 0 && first_hard[need[i]] >= 0 && first_hard[cand[i]] <= first_hard[need[i]]) break;
    bring_pos[k] = need[i];
    evict_pos[k] = cand[i];
    ++k;
  }
  *k_out = k;
  return IQSB_OK;
}
