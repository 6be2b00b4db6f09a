// kernels_fused.cu -- gate fusion: a batch of gates on low positions applied in ONE sweep
// over HBM, the tile staged in shared memory.
//
// Replaces ApplyFusedGates (reference src/qureg_fusion.cpp:55-94), which replays the queued
// gates block by block with blocks of 2^log2llc amplitudes sized for the CPU's last-level
// cache.  Here the block is a shared-memory tile of 2^K amplitudes (K = 11: 32 KiB for
// ComplexDP, several CTAs resident per SM so that the HBM traffic of one CTA overlaps the
// shared-memory sweeps of the others).  Semantics are the reference's: targets must be below
// the tile exponent; a control at or above it selects whole tiles
// (src/qureg_applyctrl1qubitgate.cpp:296-309).
#include "iqsb_internal.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kMaxFusedGates = 4096;

template <typename T>
struct FGate {
  Mat2<T> m;
  int kind, control, target, pad;
};

// 16-byte slots; XOR swizzle so that the 8 lanes of a quarter-warp hit 8 different bank
// groups for every target position ("fused kernel").
__device__ __forceinline__ unsigned phys(unsigned i) { return i ^ (((i >> 3) & 1u) * 7u); }

template <typename T>
__global__ void __launch_bounds__(kThreads)
    k_fused(Chunk<T> *__restrict__ state, uint64_t ntiles, unsigned K, const FGate<T> *__restrict__ gates, int ngates) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Cx<T> *tile = reinterpret_cast<Cx<T> *>(smem_raw);
  const unsigned nchunks = 1u << (K - 1);
  for (uint64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    Chunk<T> *g = state + (t << (K - 1));
    const uint64_t base = t << K;
    for (unsigned c = threadIdx.x; c < nchunks; c += kThreads) {
      Chunk<T> v = ld_chunk(g + c);
      tile[phys(2 * c)] = v.a;
      tile[phys(2 * c + 1)] = v.b;
    }
    __syncthreads();
    for (int gi = 0; gi < ngates; ++gi) {
      const FGate<T> G = gates[gi];
      const unsigned tp = (unsigned)G.target;
      bool controlled = G.kind == 1;
      if (controlled && (unsigned)G.control >= K) {
        if (!((base >> G.control) & 1ull)) continue;  // uniform over the CTA
        controlled = false;
      }
      if (!controlled) {
        const unsigned npairs = 1u << (K - 1);
        for (unsigned j = threadIdx.x; j < npairs; j += kThreads) {
          unsigned i0 = (unsigned)insert_zero(j, tp), i1 = i0 | (1u << tp);
          Cx<T> a = tile[phys(i0)], b = tile[phys(i1)];
          apply2x2(G.m, a, b);
          tile[phys(i0)] = a;
          tile[phys(i1)] = b;
        }
      } else {
        const unsigned cp = (unsigned)G.control;
        const unsigned lo = cp < tp ? cp : tp, hi = cp < tp ? tp : cp;
        const unsigned npairs = 1u << (K - 2);
        for (unsigned j = threadIdx.x; j < npairs; j += kThreads) {
          unsigned x = (unsigned)insert_zero(insert_zero(j, lo), hi);
          unsigned i0 = x | (1u << cp), i1 = i0 | (1u << tp);
          Cx<T> a = tile[phys(i0)], b = tile[phys(i1)];
          apply2x2(G.m, a, b);
          tile[phys(i0)] = a;
          tile[phys(i1)] = b;
        }
      }
      __syncthreads();
    }
    for (unsigned c = threadIdx.x; c < nchunks; c += kThreads) {
      Chunk<T> v;
      v.a = tile[phys(2 * c)];
      v.b = tile[phys(2 * c + 1)];
      st_chunk(g + c, v);
    }
    __syncthreads();
  }
}

template <typename T>
int run_fused(iqsb_state *st, const iqsb_fgate *gates, int ngates, unsigned K) {
  iqsb_ctx *ctx = st->ctx;
  FGate<T> *h = new FGate<T>[ngates];
  for (int i = 0; i < ngates; ++i) {
    h[i].m = make_mat<T>(gates[i].m);
    h[i].kind = gates[i].kind;
    h[i].control = gates[i].control;
    h[i].target = gates[i].target;
    h[i].pad = 0;
  }
  FGate<T> *d = nullptr;
  cudaError_t e = cudaMallocAsync((void **)&d, sizeof(FGate<T>) * ngates, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d, h, sizeof(FGate<T>) * ngates, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);  // h is pageable: copy is done after this
  delete[] h;
  if (e != cudaSuccess) {
    iqsb_set_error("iqsb_fused: staging the gate list failed: %s", cudaGetErrorString(e));
    return IQSB_ERR_CUDA;
  }
  size_t smem = (size_t)sizeof(Cx<T>) << K;
  IQSB_CUDA(cudaFuncSetAttribute(k_fused<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 1;
  IQSB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_fused<T>, kThreads, smem));
  if (per_sm < 1) per_sm = 1;
  uint64_t ntiles = st->local_amps >> K;
  uint64_t cap = (uint64_t)ctx->num_sms * per_sm;
  unsigned grid = (unsigned)(ntiles < cap ? ntiles : cap);
  k_fused<T><<<grid, kThreads, smem, ctx->stream>>>((Chunk<T> *)st->d, ntiles, K, d, ngates);
  int rc = iqsb_check_launch(ctx, "k_fused");
  cudaFreeAsync(d, ctx->stream);
  return rc;
}

}  // namespace

extern "C" int iqsb_fused_max_log2tile(const iqsb_state *st) {
  if (!st) return 0;
  int k = 11;
  if ((unsigned)k > st->log2_local) k = (int)st->log2_local;
  return k;
}

extern "C" int iqsb_fused(iqsb_state *st, const iqsb_fgate *gates, int ngates) {
  IQSB_REQUIRE(st && (gates || ngates == 0), "iqsb_fused: null argument");
  IQSB_REQUIRE(ngates >= 0 && ngates <= kMaxFusedGates, "iqsb_fused: at most %d gates per call", kMaxFusedGates);
  if (ngates == 0) return IQSB_OK;
  unsigned K = (unsigned)iqsb_fused_max_log2tile(st);
  IQSB_REQUIRE(K >= 2, "iqsb_fused: shard too small");
  for (int i = 0; i < ngates; ++i) {
    IQSB_REQUIRE(gates[i].kind == 0 || gates[i].kind == 1, "iqsb_fused: gate %d has bad kind", i);
    IQSB_REQUIRE(gates[i].target >= 0 && (unsigned)gates[i].target < K, "iqsb_fused: gate %d target %d >= tile exponent %u", i, gates[i].target, K);
    if (gates[i].kind == 1)
      IQSB_REQUIRE(gates[i].control >= 0 && (unsigned)gates[i].control < st->log2_local && gates[i].control != gates[i].target,
                   "iqsb_fused: gate %d has bad control", i);
  }
  return st->dtype == IQSB_F64 ? run_fused<double>(st, gates, ngates, K) : run_fused<float>(st, gates, ngates, K);
}
