// kernels_qaoa.cu -- the QAOA helpers of Intel-QS as device kernels.
//
// The reference keeps a classical cost function in the real part of a second "register" (diag)
// and loops over psi[i] / diag[i] on the host (src/qaoa_features.cpp):
//   InitializeVectorAs[Weighted]MaxCutCostFunction   :58-203   -> k_maxcut_int / k_maxcut_weighted
//   ImplementQaoaLayerBasedOnCostFunction            :255-266  -> k_qaoa_layer
//   GetExpectationValue[Squared]FromCostFunction     :277-335  -> k_qaoa_expect
//   GetHistogramFromCostFunction[WithWeights...]     :345-523  -> k_qaoa_hist
// All of them are one streaming pass over one or two shards: HBM-bound, same family as the
// reductions in kernels_reduce.cu.
#include <algorithm>
#include <cmath>
#include <vector>

#include "iqsb_internal.cuh"

namespace {

constexpr int kBlock = 256;
constexpr int kMaxVerts = 40;
constexpr int kMaxBins = 4096;
constexpr int kMaxTable = 2048;

struct CutGraph {
  int n;
  uint8_t pos_of_qubit[kMaxVerts];  // program qubit q sits at data position pos_of_qubit[q]
};

// program-order bit string of the global data index g
__device__ __forceinline__ uint64_t program_bits(uint64_t g, const CutGraph &G) {
  uint64_t x = 0;
  for (int q = 0; q < G.n; ++q) x |= ((g >> G.pos_of_qubit[q]) & 1ull) << q;
  return x;
}

template <typename T>
__device__ __forceinline__ void block_max_store(double v, double *partials) {
  __shared__ double sm[kBlock / 32];
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kBlock / 32; ++w) v = fmax(v, sm[w]);
    partials[blockIdx.x] = v;
  }
}

// integer adjacency: cut = (num_edges - (sum_{v,u} a[v][u] s_v s_u) / 2) / 2, exact in integers
// (reference qaoa_features.cpp:84-103)
template <typename T>
__global__ void __launch_bounds__(kBlock)
    k_maxcut_int(Cx<T> *__restrict__ diag, uint64_t n_local, uint64_t glb_start, CutGraph G, const int *__restrict__ adj, int num_edges,
                 double *partials) {
  extern __shared__ int sadj[];
  for (int k = threadIdx.x; k < G.n * G.n; k += kBlock) sadj[k] = adj[k];
  __syncthreads();
  double mx = 0;
  const uint64_t stride = (uint64_t)gridDim.x * kBlock;
  for (uint64_t i = (uint64_t)blockIdx.x * kBlock + threadIdx.x; i < n_local; i += stride) {
    uint64_t x = program_bits(glb_start + i, G);
    int sum = 0;
    for (int v = 0; v < G.n; ++v) {
      int row = 0;
      for (int u = 0; u < G.n; ++u) row += ((x >> u) & 1ull) ? sadj[v * G.n + u] : -sadj[v * G.n + u];
      sum += ((x >> v) & 1ull) ? row : -row;
    }
    int cut = (num_edges - sum / 2) / 2;
    st_amp(diag + i, Cx<T>{(T)cut, T(0)});
    mx = fmax(mx, (double)cut);
  }
  block_max_store<T>(mx, partials);
}

// weighted adjacency: the same loop order and operations as the reference (:153-160), so that the
// stored cuts are bit-identical: cut += a[v][u] * s_v * s_u ; cut = total_weight - cut/2 ; cut /= 2
template <typename T>
__global__ void __launch_bounds__(kBlock)
    k_maxcut_weighted(Cx<T> *__restrict__ diag, uint64_t n_local, uint64_t glb_start, CutGraph G, const T *__restrict__ adj, T total_weight,
                      double *partials) {
  extern __shared__ unsigned char sraw[];
  T *sadj = reinterpret_cast<T *>(sraw);
  for (int k = threadIdx.x; k < G.n * G.n; k += kBlock) sadj[k] = adj[k];
  __syncthreads();
  double mx = 0;
  const uint64_t stride = (uint64_t)gridDim.x * kBlock;
  for (uint64_t i = (uint64_t)blockIdx.x * kBlock + threadIdx.x; i < n_local; i += stride) {
    uint64_t x = program_bits(glb_start + i, G);
    T cut = 0;
    for (int v = 0; v < G.n; ++v) {
      T sv = ((x >> v) & 1ull) ? T(1) : T(-1);
      for (int u = 0; u < G.n; ++u) {
        T su = ((x >> u) & 1ull) ? T(1) : T(-1);
        cut = add_rn(cut, mul_rn(mul_rn(sadj[v * G.n + u], sv), su));
      }
    }
    cut = sub_rn(total_weight, cut / T(2));
    cut = cut / T(2);
    st_amp(diag + i, Cx<T>{cut, T(0)});
    mx = fmax(mx, (double)cut);
  }
  block_max_store<T>(mx, partials);
}

__global__ void k_max_finish(const double *partials, int n, double *out) {
  double v = 0;
  for (int b = threadIdx.x; b < n; b += 32) v = fmax(v, partials[b]);
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  if (threadIdx.x == 0) out[0] = v;
}

// psi[i] *= (cos(gamma d), -sin(gamma d)), d = Re diag[i].  For integer-valued d (unweighted MaxCut)
// the factor comes from a table the host filled with its own libm (identical bits to the reference).
template <typename T>
__global__ void __launch_bounds__(kBlock)
    k_qaoa_layer(Cx<T> *__restrict__ psi, const Cx<T> *__restrict__ diag, uint64_t n, T gamma, const Cx<T> *__restrict__ table, int ntable) {
  const uint64_t stride = (uint64_t)gridDim.x * kBlock;
  for (uint64_t i = (uint64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += stride) {
    T d = ld_amp(diag + i).re;
    Cx<T> f;
    int c = (int)d;
    if ((T)c == d && c >= 0 && c < ntable) f = table[c];
    else {
      double s, co;
      sincos((double)mul_rn(gamma, d), &s, &co);
      f = Cx<T>{(T)co, (T)(-s)};
    }
    st_amp(psi + i, cmul(ld_amp(psi + i), f));
  }
}

// out[0] = sum d |psi|^2 ; out[1] = sum d^2 |psi|^2
template <typename T>
__global__ void __launch_bounds__(kBlock)
    k_qaoa_expect(const Cx<T> *__restrict__ psi, const Cx<T> *__restrict__ diag, uint64_t n, double *partials) {
  double a0 = 0, a1 = 0;
  const uint64_t stride = (uint64_t)gridDim.x * kBlock;
  for (uint64_t i = (uint64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += stride) {
    T d = ld_amp(diag + i).re;
    T p = cnorm(ld_amp(psi + i));
    a0 += (double)mul_rn(d, p);
    a1 += (double)mul_rn(mul_rn(d, d), p);
  }
  __shared__ double sm[2][kBlock / 32];
  for (int o = 16; o > 0; o >>= 1) {
    a0 += __shfl_xor_sync(0xffffffffu, a0, o);
    a1 += __shfl_xor_sync(0xffffffffu, a1, o);
  }
  if ((threadIdx.x & 31) == 0) { sm[0][threadIdx.x >> 5] = a0; sm[1][threadIdx.x >> 5] = a1; }
  __syncthreads();
  if (threadIdx.x < 2) {
    double v = 0;
    for (int w = 0; w < kBlock / 32; ++w) v += sm[threadIdx.x][w];
    partials[(size_t)blockIdx.x * 2 + threadIdx.x] = v;
  }
}
__global__ void k_sum2_finish(const double *partials, int n, double *out) {
  double a0 = 0, a1 = 0;
  for (int b = threadIdx.x; b < n; b += 32) { a0 += partials[2 * b]; a1 += partials[2 * b + 1]; }
  for (int o = 16; o > 0; o >>= 1) {
    a0 += __shfl_xor_sync(0xffffffffu, a0, o);
    a1 += __shfl_xor_sync(0xffffffffu, a1, o);
  }
  if (threadIdx.x == 0) { out[0] = a0; out[1] = a1; }
}

// hist[floor(d / width + eps)] += |psi|^2 ; one shared-memory histogram per CTA, then global atomics
template <typename T>
__global__ void __launch_bounds__(kBlock)
    k_qaoa_hist(const Cx<T> *__restrict__ psi, const Cx<T> *__restrict__ diag, uint64_t n, int nbins, double width, double eps, double *hist,
                int *bad) {
  extern __shared__ double sh[];
  for (int k = threadIdx.x; k < nbins; k += kBlock) sh[k] = 0.0;
  __syncthreads();
  const uint64_t stride = (uint64_t)gridDim.x * kBlock;
  for (uint64_t i = (uint64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += stride) {
    double d = (double)ld_amp(diag + i).re;
    int bin = (int)floor(d / width + eps);
    if (bin < 0 || bin >= nbins) { *bad = 1; continue; }
    atomicAdd(&sh[bin], (double)cnorm(ld_amp(psi + i)));
  }
  __syncthreads();
  for (int k = threadIdx.x; k < nbins; k += kBlock)
    if (sh[k] != 0.0) atomicAdd(&hist[k], sh[k]);
}

inline int grid_for(const iqsb_ctx *ctx, uint64_t n) {
  uint64_t want = (n + kBlock - 1) / kBlock, cap = (uint64_t)ctx->num_sms * 8;
  if (cap > (uint64_t)kMaxRedBlocks) cap = kMaxRedBlocks;
  if (want < 1) want = 1;
  return (int)(want < cap ? want : cap);
}

template <typename T>
int maxcut_impl(iqsb_state *diag, unsigned nverts, const double *adjacency, int weighted, const uint8_t *pos_of_qubit, uint64_t glb_start,
                double *max_cut) {
  iqsb_ctx *ctx = diag->ctx;
  CutGraph G;
  G.n = (int)nverts;
  for (unsigned q = 0; q < nverts; ++q) G.pos_of_qubit[q] = pos_of_qubit[q];
  int grid = grid_for(ctx, diag->local_amps);
  size_t count = (size_t)nverts * nverts;
  void *d_adj = nullptr;
  int rc = IQSB_OK;
  if (!weighted) {
    std::vector<int> a(count);
    long total = 0;
    for (size_t k = 0; k < count; ++k) { a[k] = (int)adjacency[k]; total += a[k]; }
    IQSB_CUDA(cudaMalloc(&d_adj, count * sizeof(int)));
    IQSB_CUDA(cudaMemcpy(d_adj, a.data(), count * sizeof(int), cudaMemcpyHostToDevice));
    k_maxcut_int<T><<<grid, kBlock, count * sizeof(int), ctx->stream>>>((Cx<T> *)diag->d, diag->local_amps, glb_start, G, (const int *)d_adj,
                                                                         (int)(total / 2), ctx->d_partials);
  } else {
    std::vector<T> a(count);
    for (size_t k = 0; k < count; ++k) a[k] = (T)adjacency[k];
    T total_weight = 0;  // same accumulation order as the reference (:127-136)
    for (unsigned v1 = 0; v1 < nverts; ++v1)
      for (unsigned v2 = v1 + 1; v2 < nverts; ++v2)
        if (a[v1 * nverts + v2] != 0) total_weight += a[v1 * nverts + v2];
    IQSB_CUDA(cudaMalloc(&d_adj, count * sizeof(T)));
    IQSB_CUDA(cudaMemcpy(d_adj, a.data(), count * sizeof(T), cudaMemcpyHostToDevice));
    k_maxcut_weighted<T><<<grid, kBlock, count * sizeof(T), ctx->stream>>>((Cx<T> *)diag->d, diag->local_amps, glb_start, G, (const T *)d_adj,
                                                                           total_weight, ctx->d_partials);
  }
  rc = iqsb_check_launch(ctx, "k_maxcut");
  if (rc == IQSB_OK) {
    k_max_finish<<<1, 32, 0, ctx->stream>>>(ctx->d_partials, grid, ctx->d_result);
    rc = iqsb_check_launch(ctx, "k_max_finish");
  }
  cudaError_t e = cudaMemcpyAsync(ctx->h_result, ctx->d_result, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(d_adj);
  if (rc != IQSB_OK) return rc;
  if (e != cudaSuccess) {
    iqsb_set_error("iqsb_qaoa_maxcut: %s", cudaGetErrorString(e));
    return IQSB_ERR_CUDA;
  }
  *max_cut = ctx->h_result[0];
  return IQSB_OK;
}

template <typename T>
int layer_impl(iqsb_state *psi, const iqsb_state *diag, double gamma) {
  iqsb_ctx *ctx = psi->ctx;
  // table of exp(-i gamma c) for integer cuts, computed with the host's libm exactly as the reference
  // computes std::cos(gamma * diag[i].real()) / std::sin(...) (qaoa_features.cpp:264)
  std::vector<Cx<T>> table(kMaxTable);
  T g = (T)gamma;
  for (int c = 0; c < kMaxTable; ++c) {
    T arg = g * (T)c;
    table[c] = Cx<T>{(T)std::cos(arg), (T)(-std::sin(arg))};
  }
  Cx<T> *d_table = nullptr;
  IQSB_CUDA(cudaMallocAsync((void **)&d_table, sizeof(Cx<T>) * kMaxTable, ctx->stream));
  IQSB_CUDA(cudaMemcpyAsync(d_table, table.data(), sizeof(Cx<T>) * kMaxTable, cudaMemcpyHostToDevice, ctx->stream));
  IQSB_CUDA(cudaStreamSynchronize(ctx->stream));  // `table` is pageable
  int grid = (int)std::min<uint64_t>((psi->local_amps + kBlock - 1) / kBlock, (uint64_t)ctx->num_sms * 16);
  k_qaoa_layer<T><<<grid, kBlock, 0, ctx->stream>>>((Cx<T> *)psi->d, (const Cx<T> *)diag->d, psi->local_amps, g, d_table, kMaxTable);
  int rc = iqsb_check_launch(ctx, "k_qaoa_layer");
  cudaFreeAsync(d_table, ctx->stream);
  return rc;
}

}  // namespace

static bool same_shape(const iqsb_state *a, const iqsb_state *b) {
  return a && b && a->local_amps == b->local_amps && a->dtype == b->dtype && a->ctx == b->ctx;
}

extern "C" int iqsb_qaoa_maxcut(iqsb_state *diag, unsigned nverts, const double *adjacency, int weighted, const uint8_t *pos_of_qubit,
                                uint64_t glb_start, double *max_cut_local) {
  IQSB_REQUIRE(diag && adjacency && pos_of_qubit && max_cut_local, "iqsb_qaoa_maxcut: null argument");
  IQSB_REQUIRE(nverts >= 1 && nverts <= (unsigned)kMaxVerts, "iqsb_qaoa_maxcut: at most %d vertices", kMaxVerts);
  return diag->dtype == IQSB_F64 ? maxcut_impl<double>(diag, nverts, adjacency, weighted, pos_of_qubit, glb_start, max_cut_local)
                                 : maxcut_impl<float>(diag, nverts, adjacency, weighted, pos_of_qubit, glb_start, max_cut_local);
}

extern "C" int iqsb_qaoa_layer(iqsb_state *psi, const iqsb_state *diag, double gamma) {
  IQSB_REQUIRE(same_shape(psi, diag), "iqsb_qaoa_layer: registers do not match");
  return psi->dtype == IQSB_F64 ? layer_impl<double>(psi, diag, gamma) : layer_impl<float>(psi, diag, gamma);
}

extern "C" int iqsb_qaoa_expect(iqsb_state *psi, const iqsb_state *diag, double out[2]) {
  IQSB_REQUIRE(same_shape(psi, diag) && out, "iqsb_qaoa_expect: registers do not match");
  iqsb_ctx *ctx = psi->ctx;
  int grid = grid_for(ctx, psi->local_amps);
  if (psi->dtype == IQSB_F64)
    k_qaoa_expect<double><<<grid, kBlock, 0, ctx->stream>>>((const Cx<double> *)psi->d, (const Cx<double> *)diag->d, psi->local_amps, ctx->d_partials);
  else
    k_qaoa_expect<float><<<grid, kBlock, 0, ctx->stream>>>((const Cx<float> *)psi->d, (const Cx<float> *)diag->d, psi->local_amps, ctx->d_partials);
  IQSB_TRY(iqsb_check_launch(ctx, "k_qaoa_expect"));
  k_sum2_finish<<<1, 32, 0, ctx->stream>>>(ctx->d_partials, grid, ctx->d_result);
  IQSB_TRY(iqsb_check_launch(ctx, "k_sum2_finish"));
  IQSB_CUDA(cudaMemcpyAsync(ctx->h_result, ctx->d_result, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  IQSB_CUDA(cudaStreamSynchronize(ctx->stream));
  out[0] = ctx->h_result[0];
  out[1] = ctx->h_result[1];
  return IQSB_OK;
}

extern "C" int iqsb_qaoa_histogram(iqsb_state *psi, const iqsb_state *diag, int nbins, double bin_width, double eps, double *out) {
  IQSB_REQUIRE(same_shape(psi, diag) && out, "iqsb_qaoa_histogram: registers do not match");
  IQSB_REQUIRE(nbins >= 1 && nbins <= kMaxBins && bin_width > 0, "iqsb_qaoa_histogram: 1..%d bins of positive width", kMaxBins);
  iqsb_ctx *ctx = psi->ctx;
  double *d_hist = nullptr;
  IQSB_CUDA(cudaMalloc(&d_hist, sizeof(double) * nbins));
  IQSB_CUDA(cudaMemsetAsync(d_hist, 0, sizeof(double) * nbins, ctx->stream));
  IQSB_CUDA(cudaMemsetAsync(ctx->d_flags, 0, 4 * sizeof(int), ctx->stream));
  int grid = grid_for(ctx, psi->local_amps);
  if (psi->dtype == IQSB_F64)
    k_qaoa_hist<double><<<grid, kBlock, sizeof(double) * nbins, ctx->stream>>>((const Cx<double> *)psi->d, (const Cx<double> *)diag->d, psi->local_amps,
                                                                               nbins, bin_width, eps, d_hist, ctx->d_flags);
  else
    k_qaoa_hist<float><<<grid, kBlock, sizeof(double) * nbins, ctx->stream>>>((const Cx<float> *)psi->d, (const Cx<float> *)diag->d, psi->local_amps,
                                                                              nbins, bin_width, eps, d_hist, ctx->d_flags);
  int rc = iqsb_check_launch(ctx, "k_qaoa_hist");
  int bad = 0;
  cudaError_t e = cudaMemcpyAsync(out, d_hist, sizeof(double) * nbins, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(&bad, ctx->d_flags, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(d_hist);
  if (rc != IQSB_OK) return rc;
  if (e != cudaSuccess) {
    iqsb_set_error("iqsb_qaoa_histogram: %s", cudaGetErrorString(e));
    return IQSB_ERR_CUDA;
  }
  IQSB_REQUIRE(!bad, "iqsb_qaoa_histogram: a cost value falls outside [0, max_value]");
  return IQSB_OK;
}
