// exchange.cu -- exchange k local qubit positions with k global ones in ONE pass over NVLink.
//
// Reference code replaced: PermuteByLocalGlobalExchangeOfQubitPairs (src/qureg_permute.cpp:191-229)
// runs one ApplySwap per (local, global) pair, and each of those is an HP_DistrSwap
// (src/qureg_applyswap.cpp:247-480: Sendrecv into the tmp buffer, Loop_SN, Sendrecv back).  Here the
// k pairs move together: rank r keeps the 1/2^k of its shard whose local bits lpos[] already equal
// its own rank bits gpos[], and trades each of the other 2^k - 1 blocks with the partner whose rank
// bits are that block's local bits.  Per rank and direction the link carries (1 - 2^-k) * 16*L bytes
// (8*L, 12*L, 14*L for k = 1, 2, 3) against k * 8*L for k separate swaps.
//
// The move is in place: the pair (a block element here, the matching element on the partner) is owned
// by exactly one of the two ranks, chosen by one local "split" bit, and that thread loads both chunks
// and stores them crosswise -- the same ownership rule as the gate kernels of comm.cu.
//
// This is the data-movement primitive of the placement layer of the host library (src/placement.cpp):
// a global qubit that a gate needs is swapped in once and stays local for the gates that follow.
#include <string.h>

#include "iqsb_internal.cuh"

int iqsb_peer_barrier(iqsb_ctx *ctx);  // comm.cu

namespace {

constexpr int kBlock = 256;
constexpr int kUnroll = 2;
constexpr int kMaxPartners = 7;

struct XchgGeom {
  uint64_t per_partner;  // work items per partner (a power of two), in units (chunks or amplitudes)
  unsigned shift;        // log2(per_partner)
  unsigned nins;         // number of inserted zero bits (the k exchanged bits + the split bit)
  unsigned ins[4];       // ascending unit-index bit positions
  int npartners;
  uint64_t mine[kMaxPartners];    // pattern of the exchanged local bits selecting MY block for partner p
  uint64_t theirs[kMaxPartners];  // pattern selecting the partner's block
  uint64_t split[kMaxPartners];   // the split bit, set to the value this rank owns for partner p (unit index)
  void *peer[kMaxPartners];       // partner's shard (peer-mapped)
};

__device__ __forceinline__ Chunk<double> ld_unit(const Chunk<double> *p) { return ld_chunk(p); }
__device__ __forceinline__ Chunk<float> ld_unit(const Chunk<float> *p) { return ld_chunk(p); }
__device__ __forceinline__ Cx<double> ld_unit(const Cx<double> *p) { return ld_amp(p); }
__device__ __forceinline__ Cx<float> ld_unit(const Cx<float> *p) { return ld_amp(p); }
__device__ __forceinline__ void st_unit(Chunk<double> *p, const Chunk<double> &v) { st_chunk(p, v); }
__device__ __forceinline__ void st_unit(Chunk<float> *p, const Chunk<float> &v) { st_chunk(p, v); }
__device__ __forceinline__ void st_unit(Cx<double> *p, const Cx<double> &v) { st_amp(p, v); }
__device__ __forceinline__ void st_unit(Cx<float> *p, const Cx<float> &v) { st_amp(p, v); }

template <typename U>
__global__ void __launch_bounds__(kBlock) k_exchange(U *__restrict__ mine, XchgGeom g) {
  const uint64_t total = g.per_partner * (uint64_t)g.npartners;
  const uint64_t t0 = ((uint64_t)blockIdx.x * kUnroll) * kBlock + threadIdx.x;
  U a[kUnroll], b[kUnroll];
  U *pa[kUnroll], *pb[kUnroll];
#pragma unroll
  for (int u = 0; u < kUnroll; ++u) {
    const uint64_t t = t0 + (uint64_t)u * kBlock;
    if (t < total) {
      const unsigned p = (unsigned)(t >> g.shift);
      uint64_t x = t & (g.per_partner - 1);
#pragma unroll 1
      for (unsigned i = 0; i < g.nins; ++i) x = insert_zero(x, g.ins[i]);
      x |= g.split[p];
      pa[u] = mine + (x | g.mine[p]);
      pb[u] = reinterpret_cast<U *>(g.peer[p]) + (x | g.theirs[p]);
      a[u] = ld_unit(pa[u]);
      b[u] = ld_unit(pb[u]);
    }
  }
#pragma unroll
  for (int u = 0; u < kUnroll; ++u) {
    const uint64_t t = t0 + (uint64_t)u * kBlock;
    if (t < total) {
      st_unit(pa[u], b[u]);
      st_unit(pb[u], a[u]);
    }
  }
}

}  // namespace

// Pure host function (no GPU needed): what THIS rank moves when the local positions lpos[] are
// exchanged with the global positions gpos[].
extern "C" int iqsb_plan_exchange(int rank, int nranks, unsigned M, int k, const unsigned *lpos, const unsigned *gpos, iqsb_xplan *out) {
  IQSB_REQUIRE(out && lpos && gpos, "iqsb_plan_exchange: null argument");
  IQSB_REQUIRE(nranks > 1 && (nranks & (nranks - 1)) == 0 && rank >= 0 && rank < nranks, "iqsb_plan_exchange: bad rank/nranks");
  IQSB_REQUIRE(k >= 1 && k <= 3, "iqsb_plan_exchange: between 1 and 3 pairs per pass");
  IQSB_REQUIRE(M >= (unsigned)k + 1, "iqsb_plan_exchange: needs at least k + 1 local qubits");
  uint64_t used_l = 0, used_g = 0;
  for (int j = 0; j < k; ++j) {
    IQSB_REQUIRE(lpos[j] < M, "iqsb_plan_exchange: position %u is not local", lpos[j]);
    IQSB_REQUIRE(gpos[j] >= M && gpos[j] < 63 && (1ull << (gpos[j] - M)) < (uint64_t)nranks, "iqsb_plan_exchange: position %u is not global", gpos[j]);
    IQSB_REQUIRE(!((used_l >> lpos[j]) & 1) && !((used_g >> gpos[j]) & 1), "iqsb_plan_exchange: repeated position");
    used_l |= 1ull << lpos[j];
    used_g |= 1ull << gpos[j];
  }
  memset(out, 0, sizeof(*out));
  const uint64_t L = 1ull << M;
  int sb = -1;  // highest local bit that is not exchanged
  for (int b = (int)M - 1; b >= 0; --b)
    if (!((used_l >> b) & 1)) { sb = b; break; }
  out->split_bit = sb;
  out->npartners = (1 << k) - 1;
  out->amps_per_partner = L >> (k + 1);
  out->link_amps = (uint64_t)out->npartners * (L >> k);
  uint64_t my_pattern = 0;  // my rank bits, written at the exchanged local positions
  for (int j = 0; j < k; ++j) my_pattern |= (uint64_t)((rank >> (gpos[j] - M)) & 1) << lpos[j];
  for (int p = 1; p <= out->npartners; ++p) {
    int partner = rank;
    for (int j = 0; j < k; ++j)
      if ((p >> j) & 1) partner ^= 1 << (gpos[j] - M);
    uint64_t their_pattern = 0;
    for (int j = 0; j < k; ++j) their_pattern |= (uint64_t)((partner >> (gpos[j] - M)) & 1) << lpos[j];
    out->partner[p - 1] = partner;
    out->mine[p - 1] = their_pattern;  // my block with the partner's rank bits goes to the partner ...
    out->theirs[p - 1] = my_pattern;   // ... and its block with my rank bits comes here
    // the lower rank moves the pairs whose split bit is 1, the higher rank those with 0
    out->split_val[p - 1] = rank < partner ? 1 : 0;
  }
  return IQSB_OK;
}

extern "C" int iqsb_exchange_bits(iqsb_state *st, unsigned M, int k, const unsigned *lpos, const unsigned *gpos) {
  IQSB_REQUIRE(st, "iqsb_exchange_bits: null argument");
  iqsb_ctx *ctx = st->ctx;
  IQSB_REQUIRE(ctx->nranks > 1 && st->shared, "iqsb_exchange_bits: register is not shared across ranks");
  IQSB_REQUIRE(M == st->log2_local, "iqsb_exchange_bits: M must equal log2(local_amps)");
  iqsb_xplan pl;
  IQSB_TRY(iqsb_plan_exchange(ctx->rank, ctx->nranks, M, k, lpos, gpos, &pl));
  // chunk units (32 B / 16 B accesses) unless an exchanged bit or the split bit is position 0
  bool w1 = pl.split_bit == 0;
  for (int j = 0; j < k; ++j) w1 = w1 || lpos[j] == 0;
  const unsigned sh = w1 ? 0u : 1u;
  XchgGeom g;
  memset(&g, 0, sizeof(g));
  unsigned pos[4];
  int n = 0;
  for (int j = 0; j < k; ++j) pos[n++] = lpos[j];
  pos[n++] = (unsigned)pl.split_bit;
  for (int i = 0; i < n; ++i)
    for (int j = i + 1; j < n; ++j)
      if (pos[j] < pos[i]) { unsigned t = pos[i]; pos[i] = pos[j]; pos[j] = t; }
  g.nins = (unsigned)n;
  for (int i = 0; i < n; ++i) g.ins[i] = pos[i] - sh;
  g.per_partner = pl.amps_per_partner >> sh;
  unsigned lg = 0;
  while ((1ull << lg) < g.per_partner) ++lg;
  g.shift = lg;
  g.npartners = pl.npartners;
  for (int p = 0; p < pl.npartners; ++p) {
    g.mine[p] = pl.mine[p] >> sh;
    g.theirs[p] = pl.theirs[p] >> sh;
    g.split[p] = pl.split_val[p] ? (1ull << (pl.split_bit - sh)) : 0ull;
    g.peer[p] = st->peer_ptr[pl.partner[p]];
  }
  IQSB_TRY(iqsb_peer_barrier(ctx));  // the partners' earlier kernels are complete
  if (g.per_partner > 0) {
    const uint64_t total = g.per_partner * (uint64_t)g.npartners;
    const unsigned grid = (unsigned)((total + (uint64_t)kBlock * kUnroll - 1) / ((uint64_t)kBlock * kUnroll));
    if (st->dtype == IQSB_F64) {
      if (w1) k_exchange<Cx<double>><<<grid, kBlock, 0, ctx->stream>>>((Cx<double> *)st->d, g);
      else k_exchange<Chunk<double>><<<grid, kBlock, 0, ctx->stream>>>((Chunk<double> *)st->d, g);
    } else {
      if (w1) k_exchange<Cx<float>><<<grid, kBlock, 0, ctx->stream>>>((Cx<float> *)st->d, g);
      else k_exchange<Chunk<float>><<<grid, kBlock, 0, ctx->stream>>>((Chunk<float> *)st->d, g);
    }
    // "bytes" of this class = what the link carries per rank and direction (loads one way, stores the other)
    IQSB_TRY(iqsb_check_launch(ctx, "k_exchange", (double)pl.link_amps * st->amp_bytes()));
  }
  ctx->nvlink_bytes += pl.link_amps * st->amp_bytes();
  return iqsb_peer_barrier(ctx);  // the partners' stores into this shard are complete
}
