// kernels_misc.cu -- fills, axpy and the local qubit permutation.
//
// Reference loops replaced:
//   Initialize("base"/"++++")  src/qureg_init.cpp:238-244, 335-347; qureg_utils.cpp:173-184
//   AmplitudeWiseSum           src/qureg_utils.cpp:199-226
//   PermuteLocalQubits         src/qureg_permute.cpp:90-100 (there: a full copy + per-amplitude scatter;
//                              here: a few in-place tile phases through shared memory)
#include <stdlib.h>

#include "iqsb_internal.cuh"

namespace {

constexpr int kBlock = 256;
__host__ __device__ inline uint64_t div_up(uint64_t a, uint64_t b) { return (a + b - 1) / b; }

template <typename T>
__global__ void __launch_bounds__(kBlock) k_fill(Cx<T> *__restrict__ s, uint64_t n, Cx<T> v) {
  const uint64_t stride = (uint64_t)gridDim.x * kBlock;
  for (uint64_t i = (uint64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += stride) st_amp(s + i, v);
}

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
// U[-1, 1) with 53 random bits
__device__ __forceinline__ double u53(uint64_t r) { return (double)(r >> 11) * (2.0 / 9007199254740992.0) - 1.0; }

template <typename T>
__global__ void __launch_bounds__(kBlock)
    k_fill_random(Cx<T> *__restrict__ s, uint64_t n, uint64_t seed, uint64_t goff) {
  const uint64_t stride = (uint64_t)gridDim.x * kBlock;
  for (uint64_t i = (uint64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += stride) {
    uint64_t k = splitmix64(seed ^ splitmix64(2 * (goff + i)));
    uint64_t k2 = splitmix64(seed ^ splitmix64(2 * (goff + i) + 1));
    st_amp(s + i, Cx<T>{(T)u53(k), (T)u53(k2)});
  }
}

// a[i] += b[i] (UNIT) or a[i] += b[i] * f
template <typename T, bool UNIT>
__global__ void __launch_bounds__(kBlock)
    k_axpy(Chunk<T> *__restrict__ a, const Chunk<T> *__restrict__ b, uint64_t nchunks, Cx<T> f) {
  const uint64_t stride = (uint64_t)gridDim.x * kBlock;
  for (uint64_t t = (uint64_t)blockIdx.x * kBlock + threadIdx.x; t < nchunks; t += stride) {
    Chunk<T> x = ld_chunk(a + t), y = ld_chunk(b + t);
    if (UNIT) {
      x.a = cadd(x.a, y.a);
      x.b = cadd(x.b, y.b);
    } else {
      x.a = cadd(x.a, cmul(y.a, f));
      x.b = cadd(x.b, cmul(y.b, f));
    }
    st_chunk(a + t, x);
  }
}

// Local qubit permutation by in-place TILE phases.
// A phase permutes the amplitudes inside tiles: a tile is the set of 2^nS amplitudes whose indices
// differ only in the nS bit positions pos[0] < pos[1] < ... (all other bits fixed), and the phase
// applies one bit permutation sigma of those positions.  Because sigma leaves the other bits alone,
// a tile maps onto itself: the CTA loads it, scatters it through shared memory (slot t -> slot
// sigma(t)) and writes it back to the same addresses -- in place, one read and one write of HBM per
// phase, no scratch shard.  The lowest positions are always part of the tile, so global accesses are
// contiguous runs of 2^kRun amplitudes.  An arbitrary permutation of the M local qubits is split by
// the host into a few such phases (plan_phases below).
constexpr int kTileMax = 12;   // tile exponent (64 KiB of ComplexDP)
constexpr int kRun = 4;        // low positions always in the tile: 256-byte runs for ComplexDP

struct TilePhase {
  uint8_t pos[kTileMax];      // ascending positions forming the tile
  uint8_t dstslot[kTileMax];  // tile-local bit k moves to tile-local bit dstslot[k]
  int nS;
};

template <typename T, int kPermThreads>
__global__ void __launch_bounds__(kPermThreads) k_permute_tile(Cx<T> *__restrict__ state, uint64_t nouter, TilePhase ph) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Cx<T> *tile = reinterpret_cast<Cx<T> *>(smem_raw);
  // offset tables: global offset and destination slot of tile-local index t = lo | hi << 8
  __shared__ uint64_t g_lo[256], g_hi[16];
  __shared__ uint16_t l_lo[256], l_hi[16];
  const int nS = ph.nS;
  for (unsigned t = threadIdx.x; t < 256 + 16; t += kPermThreads) {
    unsigned v = t < 256 ? t : (t - 256) << 8;
    uint64_t go = 0;
    unsigned lo = 0;
    for (int k = 0; k < nS; ++k)
      if ((v >> k) & 1u) { go |= 1ull << ph.pos[k]; lo |= 1u << ph.dstslot[k]; }
    if (t < 256) { g_lo[t] = go; l_lo[t] = (uint16_t)lo; }
    else { g_hi[t - 256] = go; l_hi[t - 256] = (uint16_t)lo; }
  }
  __syncthreads();
  const unsigned tsize = 1u << nS;
  constexpr int U = 2048 / kPermThreads;  // independent 16-byte loads in flight per thread
  for (uint64_t o = blockIdx.x; o < nouter; o += gridDim.x) {
    uint64_t base = o;
    for (int k = 0; k < nS; ++k) base = insert_zero(base, ph.pos[k]);
    for (unsigned t0 = threadIdx.x; t0 < tsize; t0 += kPermThreads * U) {
      Cx<T> v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        unsigned t = t0 + u * kPermThreads;
        if (t < tsize) v[u] = ld_amp(state + (base | g_lo[t & 255] | g_hi[t >> 8]));
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        unsigned t = t0 + u * kPermThreads;
        if (t < tsize) tile[l_lo[t & 255] | l_hi[t >> 8]] = v[u];
      }
    }
    __syncthreads();
    for (unsigned t0 = threadIdx.x; t0 < tsize; t0 += kPermThreads * U) {
      Cx<T> v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        unsigned t = t0 + u * kPermThreads;
        if (t < tsize) v[u] = tile[t];
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        unsigned t = t0 + u * kPermThreads;
        if (t < tsize) st_amp(state + (base | g_lo[t & 255] | g_hi[t >> 8]), v[u]);
      }
    }
    __syncthreads();
  }
}

// ---- the same phase with the global side on the bulk-copy engine ------------------------------
// A tile is 2^(nS - kRun) runs of 2^kRun contiguous amplitudes (256 bytes of ComplexDP).  Each run
// is one cp.async.bulk (UBLKCP) from global into shared memory, completion counted in bytes on an
// mbarrier, and one cp.async.bulk back after the permutation -- no thread holds a global address
// or a staging register for the copy, and the permutation inside the tile goes through registers
// in place (8 amplitudes per thread, read before the barrier, written to their new slots after).
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
  unsigned done = 0;
  while (!done)
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, unsigned bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)), "l"(src_gmem),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *dst_gmem, const void *src_smem, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}

constexpr int kBulkThreads = 512;
template <typename T>
__global__ void __launch_bounds__(kBulkThreads) k_permute_tile_bulk(Cx<T> *__restrict__ state, uint64_t nouter, TilePhase ph) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Cx<T> *tile = reinterpret_cast<Cx<T> *>(smem_raw);
  __shared__ uint64_t g_run[256];  // offset (amplitudes) of run r inside the tile's footprint
  __shared__ uint16_t l_lo[256], l_hi[16];
  __shared__ __align__(8) uint64_t bar;
  const int nS = ph.nS;  // >= kRun, pos[0..kRun) == 0..kRun-1
  const unsigned nruns = 1u << (nS - kRun), tsize = 1u << nS;
  constexpr unsigned kRunBytes = (unsigned)sizeof(Cx<T>) << kRun;
  for (unsigned t = threadIdx.x; t < 256 + 16; t += kBulkThreads) {
    unsigned v = t < 256 ? t : (t - 256) << 8;
    unsigned lo = 0;
    for (int k = 0; k < nS; ++k)
      if ((v >> k) & 1u) lo |= 1u << ph.dstslot[k];
    if (t < 256) l_lo[t] = (uint16_t)lo;
    else l_hi[t - 256] = (uint16_t)lo;
    if (t < 256) {
      uint64_t go = 0;
      for (int k = kRun; k < nS; ++k)
        if ((t >> (k - kRun)) & 1u) go |= 1ull << ph.pos[k];
      g_run[t] = go;
    }
  }
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  unsigned parity = 0;
  constexpr int U = 4096 / kBulkThreads;
  for (uint64_t o = blockIdx.x; o < nouter; o += gridDim.x, parity ^= 1u) {
    uint64_t base = o;
    for (int k = 0; k < nS; ++k) base = insert_zero(base, ph.pos[k]);
    // a bulk copy is one warp-level instruction with uniform operands (UBLKCP): lane 0 of every warp
    // issues its share of the runs, so the 16 warps feed the copy engine side by side.  (The byte
    // count may be credited before or after the barrier is armed: the count is signed.)
    if (threadIdx.x == 0) mbar_expect_tx(&bar, tsize * (unsigned)sizeof(Cx<T>));
    if ((threadIdx.x & 31) == 0)
      for (unsigned r = threadIdx.x >> 5; r < nruns; r += kBulkThreads / 32) bulk_g2s(tile + ((size_t)r << kRun), state + (base | g_run[r]), kRunBytes, &bar);
    mbar_wait(&bar, parity);
    Cx<T> v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned t = threadIdx.x + u * kBulkThreads;
      if (t < tsize) v[u] = tile[t];
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned t = threadIdx.x + u * kBulkThreads;
      if (t < tsize) tile[l_lo[t & 255] | l_hi[t >> 8]] = v[u];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the copy engine
    __syncthreads();
    if ((threadIdx.x & 31) == 0) {
      for (unsigned r = threadIdx.x >> 5; r < nruns; r += kBulkThreads / 32) bulk_s2g(state + (base | g_run[r]), tile + ((size_t)r << kRun), kRunBytes);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the tile may be overwritten
    }
    __syncthreads();
  }
}

// Split "content of position b goes to position dst[b]" into tile phases.
static int plan_phases(const uint8_t *dst, unsigned nbits, TilePhase *out, int max_phases) {
  uint8_t cur[64];
  for (unsigned b = 0; b < nbits; ++b) cur[b] = dst[b];
  const unsigned K = nbits < (unsigned)kTileMax ? nbits : (unsigned)kTileMax;
  const unsigned run = nbits < (unsigned)kRun ? nbits : (unsigned)kRun;
  int nph = 0;
  for (;;) {
    bool todo = false;
    for (unsigned b = 0; b < nbits; ++b) todo = todo || cur[b] != b;
    if (!todo) break;
    if (nph == max_phases) return -1;
    bool in[64] = {false};
    unsigned cnt = 0;
    for (unsigned b = 0; b < run; ++b) { in[b] = true; ++cnt; }
    // pull in chains b -> cur[b] -> ... : first those that start inside the tile, then the others
    for (int pass = 0; pass < 2; ++pass)
      for (unsigned s0 = 0; s0 < nbits && cnt < K; ++s0) {
        if (cur[s0] == s0 || (pass == 0 && !in[s0])) continue;
        if (!in[s0] && K - cnt < 2) continue;  // a lone new bit cannot be finalised
        unsigned b = s0;
        do {
          if (!in[b]) { in[b] = true; ++cnt; }
          b = cur[b];
        } while (cnt < K && (!in[b] || (b != s0 && cur[b] != b && !in[cur[b]])));
      }
    // a full-size tile keeps every CTA busy: pad with the lowest positions not used yet (they stay put)
    for (unsigned b = 0; b < nbits && cnt < K; ++b)
      if (!in[b]) { in[b] = true; ++cnt; }
    // sigma on the tile: honour every destination that lies inside, park the rest on the free slots
    uint8_t sigma[64];
    bool used[64] = {false}, placed[64] = {false};
    for (unsigned b = 0; b < nbits; ++b)
      if (in[b] && in[cur[b]]) { sigma[b] = cur[b]; used[cur[b]] = true; placed[b] = true; }
    for (unsigned b = 0; b < nbits; ++b)  // stay in place when possible
      if (in[b] && !placed[b] && !used[b]) { sigma[b] = (uint8_t)b; used[b] = true; placed[b] = true; }
    for (unsigned b = 0, f = 0; b < nbits; ++b)
      if (in[b] && !placed[b]) {
        while (!in[f] || used[f]) ++f;
        sigma[b] = (uint8_t)f;
        used[f] = true;
        placed[b] = true;
      }
    // emit
    TilePhase &ph = out[nph++];
    ph.nS = 0;
    int slot_of[64];
    for (unsigned b = 0; b < nbits; ++b)
      if (in[b]) { slot_of[b] = ph.nS; ph.pos[ph.nS++] = (uint8_t)b; }
    for (int k = 0; k < ph.nS; ++k) ph.dstslot[k] = (uint8_t)slot_of[sigma[ph.pos[k]]];
    for (int k = ph.nS; k < kTileMax; ++k) ph.pos[k] = ph.dstslot[k] = 0;
    // the content that sat at b is now at sigma[b]
    uint8_t nxt[64];
    for (unsigned b = 0; b < nbits; ++b) nxt[b] = cur[b];
    for (unsigned b = 0; b < nbits; ++b)
      if (in[b]) nxt[sigma[b]] = cur[b];
    bool progress = false;
    for (unsigned b = 0; b < nbits; ++b) {
      progress = progress || (nxt[b] == b && cur[b] != b);
      cur[b] = nxt[b];
    }
    if (!progress) return -1;
  }
  return nph;
}

inline int stream_grid(const iqsb_ctx *ctx, uint64_t nwork) {
  uint64_t want = div_up(nwork, kBlock);
  uint64_t cap = (uint64_t)ctx->num_sms * 16;
  if (want < 1) want = 1;
  return (int)(want < cap ? want : cap);
}

}  // namespace

extern "C" int iqsb_fill_const(iqsb_state *st, double re, double im) {
  IQSB_REQUIRE(st, "iqsb_fill_const: null argument");
  iqsb_ctx *ctx = st->ctx;
  if (re == 0.0 && im == 0.0) {
    IQSB_CUDA(cudaMemsetAsync(st->d, 0, st->local_amps * st->amp_bytes(), ctx->stream));
    return IQSB_OK;
  }
  int grid = stream_grid(ctx, st->local_amps);
  if (st->dtype == IQSB_F64)
    k_fill<double><<<grid, kBlock, 0, ctx->stream>>>((Cx<double> *)st->d, st->local_amps, Cx<double>{re, im});
  else
    k_fill<float><<<grid, kBlock, 0, ctx->stream>>>((Cx<float> *)st->d, st->local_amps, Cx<float>{(float)re, (float)im});
  return iqsb_check_launch(ctx, "k_fill");
}

extern "C" int iqsb_fill_random(iqsb_state *st, uint64_t seed, uint64_t global_offset) {
  IQSB_REQUIRE(st, "iqsb_fill_random: null argument");
  iqsb_ctx *ctx = st->ctx;
  int grid = stream_grid(ctx, st->local_amps);
  if (st->dtype == IQSB_F64)
    k_fill_random<double><<<grid, kBlock, 0, ctx->stream>>>((Cx<double> *)st->d, st->local_amps, seed, global_offset);
  else
    k_fill_random<float><<<grid, kBlock, 0, ctx->stream>>>((Cx<float> *)st->d, st->local_amps, seed, global_offset);
  return iqsb_check_launch(ctx, "k_fill_random");
}

extern "C" int iqsb_axpy(iqsb_state *a, const iqsb_state *b, const double f[2]) {
  IQSB_REQUIRE(a && b && f, "iqsb_axpy: null argument");
  IQSB_REQUIRE(a->local_amps == b->local_amps && a->dtype == b->dtype && a->ctx == b->ctx,
               "iqsb_axpy: registers do not match");
  if (a->local_amps == 1) {  // the reference's default-constructed register: one amplitude, on the host
    double xr, xi, yr, yi;
    IQSB_TRY(iqsb_get_amp(a, 0, &xr, &xi));
    IQSB_TRY(iqsb_get_amp(const_cast<iqsb_state *>(b), 0, &yr, &yi));
    if (a->dtype == IQSB_F64) return iqsb_set_amp(a, 0, xr + (f[0] * yr - f[1] * yi), xi + (f[0] * yi + f[1] * yr));
    const float fr = (float)f[0], fi = (float)f[1], br = (float)yr, bi = (float)yi;
    return iqsb_set_amp(a, 0, (float)xr + (fr * br - fi * bi), (float)xi + (fr * bi + fi * br));
  }
  iqsb_ctx *ctx = a->ctx;
  uint64_t nchunks = a->local_amps / 2;
  int grid = stream_grid(ctx, nchunks);
  bool unit = (f[0] == 1.0 && f[1] == 0.0);  // reference qureg_utils.cpp:201
  if (a->dtype == IQSB_F64) {
    Cx<double> ff{f[0], f[1]};
    if (unit) k_axpy<double, true><<<grid, kBlock, 0, ctx->stream>>>((Chunk<double> *)a->d, (const Chunk<double> *)b->d, nchunks, ff);
    else k_axpy<double, false><<<grid, kBlock, 0, ctx->stream>>>((Chunk<double> *)a->d, (const Chunk<double> *)b->d, nchunks, ff);
  } else {
    Cx<float> ff{(float)f[0], (float)f[1]};
    if (unit) k_axpy<float, true><<<grid, kBlock, 0, ctx->stream>>>((Chunk<float> *)a->d, (const Chunk<float> *)b->d, nchunks, ff);
    else k_axpy<float, false><<<grid, kBlock, 0, ctx->stream>>>((Chunk<float> *)a->d, (const Chunk<float> *)b->d, nchunks, ff);
  }
  return iqsb_check_launch(ctx, "k_axpy");
}

// Pure host function (no GPU needed): the tile phases iqsb_permute_local would run.
// out[p*25 + 0] = nS, out[p*25 + 1 .. 12] = pos[], out[p*25 + 13 .. 24] = dstslot[].
extern "C" int iqsb_plan_permute(const uint8_t *dst_bit, unsigned nbits, uint8_t *out, int max_phases, int *nphases) {
  IQSB_REQUIRE(dst_bit && out && nphases && nbits <= 63 && max_phases > 0 && max_phases <= 64, "iqsb_plan_permute: bad argument");
  TilePhase phases[64];
  int nph = plan_phases(dst_bit, nbits, phases, max_phases);
  IQSB_REQUIRE(nph >= 0, "iqsb_plan_permute: planning failed");
  for (int p = 0; p < nph; ++p) {
    out[p * 25] = (uint8_t)phases[p].nS;
    for (int k = 0; k < kTileMax; ++k) {
      out[p * 25 + 1 + k] = phases[p].pos[k];
      out[p * 25 + 13 + k] = phases[p].dstslot[k];
    }
  }
  *nphases = nph;
  return IQSB_OK;
}

// new[j] = old[i], bit b of i -> bit dst_bit[b] of j.
extern "C" int iqsb_permute_local(iqsb_state *st, const uint8_t *dst_bit, unsigned nbits) {
  IQSB_REQUIRE(st && dst_bit, "iqsb_permute_local: null argument");
  IQSB_REQUIRE(nbits == st->log2_local, "iqsb_permute_local: nbits must equal log2(local_amps)");
  uint64_t seen = 0;
  bool identity = true;
  IQSB_REQUIRE(nbits <= 63, "iqsb_permute_local: too many bits");
  for (unsigned b = 0; b < nbits; ++b) {
    IQSB_REQUIRE(dst_bit[b] < nbits && !((seen >> dst_bit[b]) & 1), "iqsb_permute_local: not a permutation");
    seen |= 1ull << dst_bit[b];
    identity = identity && dst_bit[b] == b;
  }
  if (identity) return IQSB_OK;
  iqsb_ctx *ctx = st->ctx;
  TilePhase phases[64];
  int nph = plan_phases(dst_bit, nbits, phases, 64);
  IQSB_REQUIRE(nph >= 0, "iqsb_permute_local: internal error while planning the tile phases");
  static int threads = 0;
  if (!threads) {
    const char *e = getenv("IQSB_PERM_THREADS");
    threads = e ? atoi(e) : 512;
    if (threads != 256 && threads != 512 && threads != 1024) threads = 512;
  }
  // IQS_B200_PERMUTE_BULK=1 moves full tiles with the bulk-copy engine (k_permute_tile_bulk).  Measured at 32
  // qubits (profiles/r02o_permute_bulk_n32.log): 27.2 ms per phase against 25.0 ms for the per-thread
  // 16-byte accesses -- a tile built from arbitrary positions only offers 256-byte runs, too short for
  // the engine, and the register pass needs 2 CTAs of 512 threads per SM -- so it stays opt-in.
  const char *be = getenv("IQS_B200_PERMUTE_BULK");
  const bool use_bulk = be && *be == '1';
  for (int p = 0; p < nph; ++p) {
    const TilePhase &ph = phases[p];
    size_t smem = st->amp_bytes() << ph.nS;
    uint64_t nouter = st->local_amps >> ph.nS;
#define IQSB_PERM_LAUNCH(T, TH)                                                                                        \
  {                                                                                                                    \
    int per_sm = 1;                                                                                                    \
    IQSB_CUDA(cudaFuncSetAttribute(k_permute_tile<T, TH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
    IQSB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_permute_tile<T, TH>, TH, smem));                \
    if (per_sm < 1) per_sm = 1;                                                                                        \
    uint64_t cap = (uint64_t)ctx->num_sms * per_sm;                                                                    \
    unsigned grid = (unsigned)(nouter < cap ? nouter : cap);                                                           \
    k_permute_tile<T, TH><<<grid, TH, smem, ctx->stream>>>((Cx<T> *)st->d, nouter, ph);                                \
  }
#define IQSB_PERM_BULK(T)                                                                                              \
  {                                                                                                                    \
    int per_sm = 1;                                                                                                    \
    IQSB_CUDA(cudaFuncSetAttribute(k_permute_tile_bulk<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   \
    IQSB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_permute_tile_bulk<T>, kBulkThreads, smem));     \
    if (per_sm < 1) per_sm = 1;                                                                                        \
    uint64_t cap = (uint64_t)ctx->num_sms * per_sm;                                                                    \
    unsigned grid = (unsigned)(nouter < cap ? nouter : cap);                                                           \
    k_permute_tile_bulk<T><<<grid, kBulkThreads, smem, ctx->stream>>>((Cx<T> *)st->d, nouter, ph);                      \
  }
    // full-size tiles whose lowest kRun slots are the lowest kRun positions move through the bulk-copy engine
    bool bulk = use_bulk && ph.nS == kTileMax;
    for (int k = 0; k < kRun && bulk; ++k) bulk = ph.pos[k] == k;
    if (bulk) {
      if (st->dtype == IQSB_F64) IQSB_PERM_BULK(double) else IQSB_PERM_BULK(float)
    } else if (st->dtype == IQSB_F64) {
      if (threads == 256) IQSB_PERM_LAUNCH(double, 256) else if (threads == 1024) IQSB_PERM_LAUNCH(double, 1024) else IQSB_PERM_LAUNCH(double, 512)
    } else {
      if (threads == 256) IQSB_PERM_LAUNCH(float, 256) else if (threads == 1024) IQSB_PERM_LAUNCH(float, 1024) else IQSB_PERM_LAUNCH(float, 512)
    }
#undef IQSB_PERM_LAUNCH
#undef IQSB_PERM_BULK
    IQSB_TRY(iqsb_check_launch(ctx, bulk ? "k_permute_tile_bulk" : "k_permute_tile", 2.0 * (double)st->local_amps * st->amp_bytes()));
  }
  return IQSB_OK;
}
