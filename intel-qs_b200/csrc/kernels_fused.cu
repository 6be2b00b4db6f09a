// kernels_fused.cu -- gate fusion: a batch of gates applied in as few sweeps over HBM as
// possible, the amplitudes staged in shared-memory tiles.
//
// Replaces ApplyFusedGates (reference src/qureg_fusion.cpp:55-94), which replays the queued gates
// block by block, with blocks of 2^log2llc CONTIGUOUS amplitudes sized for the CPU's last-level
// cache -- so only gates whose target lies below log2llc can be fused there.
//
// Here a tile is the set of 2^K amplitudes (K = 12: 64 KiB of ComplexDP) whose indices differ only
// in K chosen bit positions pos[0] < pos[1] < ...: the four lowest positions are always part of it
// (global accesses are 256-byte runs, moved as 32-byte chunks), the other eight are whatever
// positions the gates of the run act on.  A run of consecutive gates whose targets fit in one tile
// costs ONE read and ONE write of the state, wherever the target qubits sit; iqsb_fused cuts a
// batch into such runs greedily.  Controls may be anywhere: inside the tile they are a bit of the
// tile-local index, outside they select whole tiles (the reference's rule for controls above the
// block, src/qureg_applyctrl1qubitgate.cpp:296-309).
//
// Inside a tile every gate is one shared-memory round trip of the tile (16-byte slots, XOR-swizzled
// so that the 8 lanes of a quarter-warp hit 8 different bank groups for every target slot).
// Measured (profiles/r01_ncu_summary.md): a run costs one sweep (78-91 % of the copy peak) plus
// 1.36 ms per gate per 2^30 amplitudes inside the tile, 3.5x cheaper than a sweep per gate; that
// in-tile cost is bound by the FP64 issue rate of the exact, non-contracted arithmetic (28
// instructions per pair), 1.05 ms with IQSB_ARITH_FMA (16 instructions, then shared-memory bound).
#include <string.h>

#include <vector>

#include "iqsb_internal.cuh"

namespace {

// Tuning knobs (compile-time; the defaults are the measured best, profiles/r01_fused_variants.md)
#ifndef IQSB_FUSED_THREADS
#define IQSB_FUSED_THREADS 256
#endif
#ifndef IQSB_FUSED_MINBLOCKS
#define IQSB_FUSED_MINBLOCKS 4
#endif
#ifndef IQSB_FUSED_TILE
#define IQSB_FUSED_TILE 12
#endif
#ifndef IQSB_FUSED_LOW
#define IQSB_FUSED_LOW 4
#endif
#ifndef IQSB_FUSED_LOADS
#define IQSB_FUSED_LOADS 2
#endif
#ifndef IQSB_FUSED_ASYNC_LOAD
#define IQSB_FUSED_ASYNC_LOAD 1
#endif
#ifndef IQSB_FUSED_PAIR_UNROLL
#define IQSB_FUSED_PAIR_UNROLL 2
#endif
constexpr int kThreads = IQSB_FUSED_THREADS;
constexpr int kMaxFusedGates = 4096;
constexpr int kTile = IQSB_FUSED_TILE;  // tile exponent (<= 12)
constexpr int kLow = IQSB_FUSED_LOW;    // lowest positions always in the tile
static_assert(kTile >= 9 && kTile <= 11 + 1 && kLow >= 1 && kLow <= 4, "tile geometry");
constexpr int kPairUnroll = IQSB_FUSED_PAIR_UNROLL;

template <typename T>
struct alignas(16) FGate {
  Mat2<T> m;
  int tslot;  // tile-local bit of the target
  int ckind;  // 0: none, 1: control is tile-local bit `c`, 2: control is bit `c` of the global index (outside the tile)
  int c;
  int pad;
};

struct TileDesc {
  uint8_t pos[kTile];
  int nS;
};

// 16-byte slots; swizzle so that pairs (i, i + 2^s) are conflict free for every s
__device__ __forceinline__ unsigned phys(unsigned i) { return i ^ (((i >> 3) & 1u) * 7u); }

// global -> tile without register staging: one asynchronous 16-byte (ComplexDP) / 8-byte (ComplexSP)
// copy per amplitude, straight into its swizzled slot (LDGSTS).  Every thread has its whole share
// of the tile in flight at once, which is what brings the sweep of a run to the copy bandwidth
// (profiles/r01_ncu_summary.md: 8 x 32 B per thread in one batch = 101 % of the copy peak, but 64
// staging registers cost a CTA per SM; the asynchronous copies need none).
__device__ __forceinline__ void cp_async_amp(Cx<double> *smem, const Cx<double> *gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_amp(Cx<float> *smem, const Cx<float> *gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gmem) : "memory");
}
template <typename T>
__device__ __forceinline__ void tile_load_async(Cx<T> *tile, const Chunk<T> *g, const uint64_t *g_lo, const uint64_t *g_hi, unsigned nchunks) {
#pragma unroll 4
  for (unsigned c = threadIdx.x; c < nchunks; c += kThreads) {
    const Cx<T> *src = reinterpret_cast<const Cx<T> *>(g + (g_lo[c & 255] | g_hi[c >> 8]));
    cp_async_amp(tile + phys(2 * c), src);
    cp_async_amp(tile + phys(2 * c + 1), src + 1);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// global <-> tile, U 32-byte accesses in flight per thread; nchunks is a multiple of kThreads * U
template <typename T, int U>
__device__ __forceinline__ void tile_load(Cx<T> *tile, const Chunk<T> *g, const uint64_t *g_lo, const uint64_t *g_hi, unsigned nchunks) {
#pragma unroll 1
  for (unsigned c0 = threadIdx.x; c0 < nchunks; c0 += kThreads * U) {
    Chunk<T> v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      unsigned c = c0 + u * kThreads;
      v[u] = ld_chunk(g + (g_lo[c & 255] | g_hi[c >> 8]));
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      unsigned c = c0 + u * kThreads;
      tile[phys(2 * c)] = v[u].a;
      tile[phys(2 * c + 1)] = v[u].b;
    }
  }
}
template <typename T, int U>
__device__ __forceinline__ void tile_store(const Cx<T> *tile, Chunk<T> *g, const uint64_t *g_lo, const uint64_t *g_hi, unsigned nchunks) {
#pragma unroll 1
  for (unsigned c0 = threadIdx.x; c0 < nchunks; c0 += kThreads * U) {
    Chunk<T> v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      unsigned c = c0 + u * kThreads;
      v[u].a = tile[phys(2 * c)];
      v[u].b = tile[phys(2 * c + 1)];
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      unsigned c = c0 + u * kThreads;
      st_chunk(g + (g_lo[c & 255] | g_hi[c >> 8]), v[u]);
    }
  }
}

constexpr int kGateBatch = 16;  // gate descriptors staged in shared memory at a time

template <typename T, bool FMA>
__global__ void __launch_bounds__(kThreads, IQSB_FUSED_MINBLOCKS)
    k_fused(Chunk<T> *__restrict__ state, uint64_t nouter, TileDesc td, const FGate<T> *__restrict__ gates, int ngates) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Cx<T> *tile = reinterpret_cast<Cx<T> *>(smem_raw);
  // chunk offset of tile-local chunk index c = lo | hi << 8 (c = tile-local amplitude index / 2)
  __shared__ uint64_t g_lo[256], g_hi[8];
  __shared__ __align__(16) FGate<T> s_gate[kGateBatch];
  __shared__ int s_pos[16];
  const int nS = td.nS;  // pos[0] == 0 always
  if (threadIdx.x < kTile) s_pos[threadIdx.x] = td.pos[threadIdx.x];
  __syncthreads();
  for (unsigned t = threadIdx.x; t < 256 + 8; t += kThreads) {
    unsigned v = t < 256 ? t : (t - 256) << 8;
    uint64_t go = 0;
#pragma unroll 1
    for (int k = 1; k < nS; ++k)
      if ((v >> (k - 1)) & 1u) go |= 1ull << (s_pos[k] - 1);
    if (t < 256) g_lo[t] = go;
    else g_hi[t - 256] = go;
  }
  __syncthreads();
  const unsigned nchunks = 1u << (nS - 1);
  constexpr int U = IQSB_FUSED_LOADS;  // 32-byte loads in flight per thread
  for (uint64_t o = blockIdx.x; o < nouter; o += gridDim.x) {
    uint64_t base = o;  // amplitude index with zeros at the tile positions
#pragma unroll 1
    for (int k = 0; k < nS; ++k) base = insert_zero(base, (unsigned)s_pos[k]);
    Chunk<T> *g = state + (base >> 1);
#if IQSB_FUSED_ASYNC_LOAD
    tile_load_async<T>(tile, g, g_lo, g_hi, nchunks);
#else
    if (nchunks % (kThreads * U) == 0) tile_load<T, U>(tile, g, g_lo, g_hi, nchunks);
    else tile_load<T, 1>(tile, g, g_lo, g_hi, nchunks);
#endif
    for (int g0 = 0; g0 < ngates; g0 += kGateBatch) {
      // stage the next descriptors (the barrier also orders the tile accesses of the previous gate)
      const int nb = ngates - g0 < kGateBatch ? ngates - g0 : kGateBatch;
      constexpr int kWords = (int)(sizeof(FGate<T>) / 16);
      if (g0) __syncthreads();  // a skipped gate has no barrier of its own: nobody still reads s_gate
      if ((int)threadIdx.x < nb * kWords)
        reinterpret_cast<int4 *>(s_gate)[threadIdx.x] = __ldg(reinterpret_cast<const int4 *>(gates + g0) + threadIdx.x);
      __syncthreads();
      for (int gi = 0; gi < nb; ++gi) {
        const unsigned ts = (unsigned)s_gate[gi].tslot;
        const int ckind = s_gate[gi].ckind;
        const unsigned cs = (unsigned)s_gate[gi].c;
        if (ckind == 2 && !((base >> cs) & 1ull)) continue;  // uniform over the CTA
        const Mat2<T> m = s_gate[gi].m;
        if (ckind != 1) {
          const unsigned npairs = 1u << (nS - 1);
#pragma unroll kPairUnroll
          for (unsigned j = threadIdx.x; j < npairs; j += kThreads) {
            const unsigned x = (unsigned)insert_zero(j, ts);
            const unsigned i0 = phys(x), i1 = phys(x | (1u << ts));
            Cx<T> a = tile[i0], b = tile[i1];
            if (FMA) apply2x2_fma(m, a, b);
            else apply2x2(m, a, b);
            tile[i0] = a;
            tile[i1] = b;
          }
        } else {
          const unsigned lo = cs < ts ? cs : ts, hi = cs < ts ? ts : cs;
          const unsigned npairs = 1u << (nS - 2);
#pragma unroll kPairUnroll
          for (unsigned j = threadIdx.x; j < npairs; j += kThreads) {
            const unsigned x = (unsigned)insert_zero(insert_zero(j, lo), hi) | (1u << cs);
            const unsigned i0 = phys(x), i1 = phys(x | (1u << ts));
            Cx<T> a = tile[i0], b = tile[i1];
            if (FMA) apply2x2_fma(m, a, b);
            else apply2x2(m, a, b);
            tile[i0] = a;
            tile[i1] = b;
          }
        }
        __syncthreads();
      }
    }
    __syncthreads();  // covers ngates == 0 and a skipped last gate
    if (nchunks % (kThreads * U) == 0) tile_store<T, U>(tile, g, g_lo, g_hi, nchunks);
    else tile_store<T, 1>(tile, g, g_lo, g_hi, nchunks);
    __syncthreads();
  }
}

// one run: gates [first, last) of `in` all have their target in the tile `td`
template <typename T>
int launch_run(iqsb_state *st, const iqsb_fgate *in, int first, int last, const TileDesc &td) {
  iqsb_ctx *ctx = st->ctx;
  int slot_of[64];
  for (int p = 0; p < 64; ++p) slot_of[p] = -1;
  for (int k = 0; k < td.nS; ++k) slot_of[td.pos[k]] = k;
  std::vector<FGate<T>> gates((size_t)(last - first));
  for (int k = first; k < last; ++k) {
    FGate<T> &o = gates[k - first];
    o.m = make_mat<T>(in[k].m);
    o.tslot = slot_of[in[k].target];
    o.pad = 0;
    o.ckind = 0;
    o.c = 0;
    if (in[k].kind == 1) {
      int c = in[k].control;
      if (slot_of[c] >= 0) { o.ckind = 1; o.c = slot_of[c]; }
      else { o.ckind = 2; o.c = c; }
    }
  }
  // descriptors travel through the context's staging ring: written into pinned memory, copied in
  // stream order; the host only waits when the ring wraps around
  const size_t bytes = sizeof(FGate<T>) * gates.size();
  if (!ctx->stage_h) {
    IQSB_CUDA(cudaMallocHost((void **)&ctx->stage_h, kStageBytes));
    IQSB_CUDA(cudaMalloc((void **)&ctx->stage_d, kStageBytes));
    ctx->stage_off = 0;
  }
  IQSB_REQUIRE(bytes <= kStageBytes, "iqsb_fused: gate list too long for the staging ring");
  if (ctx->stage_off + bytes > kStageBytes) {
    IQSB_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->stage_off = 0;
  }
  memcpy(ctx->stage_h + ctx->stage_off, gates.data(), bytes);
  FGate<T> *d = reinterpret_cast<FGate<T> *>(ctx->stage_d + ctx->stage_off);
  IQSB_CUDA(cudaMemcpyAsync(d, ctx->stage_h + ctx->stage_off, bytes, cudaMemcpyHostToDevice, ctx->stream));
  ctx->stage_off += (bytes + 255) & ~(size_t)255;
  size_t smem = (size_t)sizeof(Cx<T>) << td.nS;
  auto kernel = ctx->arith == IQSB_ARITH_FMA ? k_fused<T, true> : k_fused<T, false>;
  IQSB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 1;
  IQSB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, smem));
  if (per_sm < 1) per_sm = 1;
  uint64_t nouter = st->local_amps >> td.nS;
  uint64_t cap = (uint64_t)ctx->num_sms * per_sm;
  unsigned grid = (unsigned)(nouter < cap ? nouter : cap);
  kernel<<<grid, kThreads, smem, ctx->stream>>>((Chunk<T> *)st->d, nouter, td, d, (int)gates.size());
  return iqsb_check_launch(ctx, "k_fused");
}

}  // namespace

extern "C" int iqsb_fused_max_log2tile(const iqsb_state *st) {
  if (!st) return 0;
  int k = kTile;
  if ((unsigned)k > st->log2_local) k = (int)st->log2_local;
  return k;
}

// Pure host function: how iqsb_fused cuts a batch into runs.  run_end[r] = index one past the last
// gate of run r; tiles[r*16] = number of tile positions, tiles[r*16 + 1 ..] = the positions.
extern "C" int iqsb_plan_fused(const iqsb_fgate *gates, int ngates, unsigned log2_local, int *run_end, uint8_t *tiles, int max_runs, int *nruns) {
  IQSB_REQUIRE((gates || ngates == 0) && run_end && tiles && nruns && max_runs > 0, "iqsb_plan_fused: null argument");
  const unsigned K = log2_local < (unsigned)kTile ? log2_local : (unsigned)kTile;
  const unsigned low = log2_local < (unsigned)kLow ? log2_local : (unsigned)kLow;
  int r = 0, first = 0;
  while (first < ngates) {
    IQSB_REQUIRE(r < max_runs, "iqsb_plan_fused: more than %d runs", max_runs);
    bool in[64] = {false};
    unsigned cnt = 0;
    for (unsigned b = 0; b < low; ++b) { in[b] = true; ++cnt; }
    int last = first;
    for (; last < ngates; ++last) {
      unsigned t = (unsigned)gates[last].target;
      IQSB_REQUIRE(t < log2_local, "iqsb_fused: gate %d: target %u is not a local position", last, t);
      if (!in[t]) {
        if (cnt == K) break;
        in[t] = true;
        ++cnt;
      }
    }
    // use the spare slots for controls of the run (cheaper inside the tile), then for low positions
    for (int k = first; k < last && cnt < K; ++k)
      if (gates[k].kind == 1 && (unsigned)gates[k].control < log2_local && !in[gates[k].control]) { in[gates[k].control] = true; ++cnt; }
    for (unsigned b = 0; b < log2_local && cnt < K; ++b)
      if (!in[b]) { in[b] = true; ++cnt; }
    uint8_t *td = tiles + r * 16;
    td[0] = (uint8_t)cnt;
    int n = 0;
    for (unsigned b = 0; b < log2_local; ++b)
      if (in[b]) td[1 + n++] = (uint8_t)b;
    for (; n < 15; ++n) td[1 + n] = 0;
    run_end[r++] = last;
    first = last;
  }
  *nruns = r;
  return IQSB_OK;
}

extern "C" int iqsb_fused(iqsb_state *st, const iqsb_fgate *gates, int ngates) {
  IQSB_REQUIRE(st && (gates || ngates == 0), "iqsb_fused: null argument");
  IQSB_REQUIRE(ngates >= 0 && ngates <= kMaxFusedGates, "iqsb_fused: at most %d gates per call", kMaxFusedGates);
  if (ngates == 0) return IQSB_OK;
  for (int i = 0; i < ngates; ++i) {
    IQSB_REQUIRE(gates[i].kind == 0 || gates[i].kind == 1, "iqsb_fused: gate %d has bad kind", i);
    IQSB_REQUIRE(gates[i].target >= 0 && (unsigned)gates[i].target < st->log2_local, "iqsb_fused: gate %d: target %d is not a local position", i,
                 gates[i].target);
    if (gates[i].kind == 1)
      IQSB_REQUIRE(gates[i].control >= 0 && (unsigned)gates[i].control < st->log2_local && gates[i].control != gates[i].target,
                   "iqsb_fused: gate %d has bad control", i);
  }
  if (st->log2_local < 2) {  // nothing to tile
    for (int i = 0; i < ngates; ++i) {
      if (gates[i].kind == 0) IQSB_TRY(iqsb_gate1(st, (unsigned)gates[i].target, gates[i].m, 0, st->local_amps));
      else IQSB_TRY(iqsb_cgate1(st, (unsigned)gates[i].control, (unsigned)gates[i].target, gates[i].m, 0, st->local_amps));
    }
    return IQSB_OK;
  }
  std::vector<int> run_end((size_t)ngates);
  std::vector<uint8_t> tiles((size_t)ngates * 16);
  int nruns = 0;
  IQSB_TRY(iqsb_plan_fused(gates, ngates, st->log2_local, run_end.data(), tiles.data(), ngates, &nruns));
  int first = 0;
  for (int r = 0; r < nruns; ++r) {
    TileDesc td;
    td.nS = tiles[r * 16];
    for (int k = 0; k < kTile; ++k) td.pos[k] = tiles[r * 16 + 1 + k];
    int rc = st->dtype == IQSB_F64 ? launch_run<double>(st, gates, first, run_end[r], td) : launch_run<float>(st, gates, first, run_end[r], td);
    if (rc != IQSB_OK) return rc;
    first = run_end[r];
  }
  return IQSB_OK;
}
