// kernels_misc.cu -- fills, axpy and the local qubit permutation.
//
// Reference loops replaced:
//   Initialize("base"/"++++")  src/qureg_init.cpp:238-244, 335-347; qureg_utils.cpp:173-184
//   AmplitudeWiseSum           src/qureg_utils.cpp:199-226
//   PermuteLocalQubits         src/qureg_permute.cpp:90-100 (out-of-place bit permutation)
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

// Bit permutation, gather form: dst[j] = src[pext-like(j)].  src_bit[b] = which bit of the
// SOURCE index supplies bit b of the destination index... we need the inverse: for output j,
// source i has bit sb = bit dst_of[sb] of j.  The table `src_from[b]` gives, for source bit b,
// the destination bit that holds it.  Bits below `keep` are untouched (identity), so a thread
// moves a contiguous run of 2^keep amplitudes when keep >= 1.
struct PermTable {
  uint8_t dst_bit[64];
};

template <typename T>
__global__ void __launch_bounds__(kBlock)
    k_permute_gather_w2(const Chunk<T> *__restrict__ src, Chunk<T> *__restrict__ dst, uint64_t nchunks,
                        unsigned nbits, PermTable tab) {
  // chunk index c covers amplitude bits 1..nbits-1; bit 0 is fixed (dst_bit[0] == 0).
  const uint64_t stride = (uint64_t)gridDim.x * kBlock;
  for (uint64_t j = (uint64_t)blockIdx.x * kBlock + threadIdx.x; j < nchunks; j += stride) {
    uint64_t i = 0;
    for (unsigned b = 1; b < nbits; ++b) i |= ((j >> (tab.dst_bit[b] - 1)) & 1ull) << (b - 1);
    st_chunk(dst + j, ld_chunk(src + i));
  }
}
template <typename T>
__global__ void __launch_bounds__(kBlock)
    k_permute_gather_w1(const Cx<T> *__restrict__ src, Cx<T> *__restrict__ dst, uint64_t n, unsigned nbits,
                        PermTable tab) {
  const uint64_t stride = (uint64_t)gridDim.x * kBlock;
  for (uint64_t j = (uint64_t)blockIdx.x * kBlock + threadIdx.x; j < n; j += stride) {
    uint64_t i = 0;
    for (unsigned b = 0; b < nbits; ++b) i |= ((j >> tab.dst_bit[b]) & 1ull) << b;
    st_amp(dst + j, ld_amp(src + i));
  }
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
  IQSB_REQUIRE(a->local_amps >= 2, "iqsb_axpy: shard too small");
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

// new[j] = old[i], bit b of i -> bit dst_bit[b] of j.
extern "C" int iqsb_permute_local(iqsb_state *st, const uint8_t *dst_bit, unsigned nbits) {
  IQSB_REQUIRE(st && dst_bit, "iqsb_permute_local: null argument");
  IQSB_REQUIRE(nbits == st->log2_local, "iqsb_permute_local: nbits must equal log2(local_amps)");
  uint64_t seen = 0;
  bool identity = true;
  PermTable tab;
  for (unsigned b = 0; b < 64; ++b) tab.dst_bit[b] = (uint8_t)b;
  for (unsigned b = 0; b < nbits; ++b) {
    IQSB_REQUIRE(dst_bit[b] < nbits && !((seen >> dst_bit[b]) & 1), "iqsb_permute_local: not a permutation");
    seen |= 1ull << dst_bit[b];
    tab.dst_bit[b] = dst_bit[b];
    identity = identity && dst_bit[b] == b;
  }
  if (identity) return IQSB_OK;
  iqsb_ctx *ctx = st->ctx;
  size_t bytes = st->local_amps * st->amp_bytes();
  // A bit permutation is a product of (nbits - #cycles) transpositions, and a transposition is a
  // SWAP sweep that moves half of the shard in place (16*L bytes, no scratch).  The out-of-place
  // gather below costs a read of scattered 16/32-byte pieces plus two more passes for the copy back
  // and needs a second shard of HBM; measured at 2^32 amplitudes a full bit reversal takes 250 ms that
  // way against ~10 ms per transposition.  So: transpositions unless there are very many of them.
  unsigned transpositions = 0;
  {
    uint64_t visited = 0;
    for (unsigned b = 0; b < nbits; ++b) {
      if ((visited >> b) & 1) continue;
      unsigned len = 0;
      for (unsigned c = b; !((visited >> c) & 1); c = dst_bit[c]) { visited |= 1ull << c; ++len; }
      transpositions += len - 1;
    }
  }
  void *scratch = nullptr;
  cudaError_t e = transpositions <= 24 ? cudaErrorMemoryAllocation : cudaMalloc(&scratch, bytes);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    uint8_t cur[64];  // cur[b] = destination still owed by the data sitting at source bit b
    for (unsigned b = 0; b < nbits; ++b) cur[b] = dst_bit[b];
    const double X[8] = {0, 0, 1, 0, 1, 0, 0, 0};
    for (unsigned b = 0; b < nbits; ++b) {
      while (cur[b] != b) {
        unsigned t = cur[b];  // content of bit b must go to bit t: swap bits b and t
        IQSB_TRY(iqsb_swap2x2(st, b < t ? b : t, b < t ? t : b, X));
        uint8_t tmp = cur[t];
        cur[t] = (uint8_t)t;  // bit t now holds its final content
        cur[b] = tmp;
      }
    }
    return IQSB_OK;
  }
  int rc = IQSB_OK;
  if (dst_bit[0] == 0 && st->local_amps >= 2) {
    uint64_t nchunks = st->local_amps / 2;
    int grid = stream_grid(ctx, nchunks);
    if (st->dtype == IQSB_F64)
      k_permute_gather_w2<double><<<grid, kBlock, 0, ctx->stream>>>((const Chunk<double> *)st->d, (Chunk<double> *)scratch, nchunks, nbits, tab);
    else
      k_permute_gather_w2<float><<<grid, kBlock, 0, ctx->stream>>>((const Chunk<float> *)st->d, (Chunk<float> *)scratch, nchunks, nbits, tab);
  } else {
    int grid = stream_grid(ctx, st->local_amps);
    if (st->dtype == IQSB_F64)
      k_permute_gather_w1<double><<<grid, kBlock, 0, ctx->stream>>>((const Cx<double> *)st->d, (Cx<double> *)scratch, st->local_amps, nbits, tab);
    else
      k_permute_gather_w1<float><<<grid, kBlock, 0, ctx->stream>>>((const Cx<float> *)st->d, (Cx<float> *)scratch, st->local_amps, nbits, tab);
  }
  rc = iqsb_check_launch(ctx, "k_permute_gather");
  if (rc == IQSB_OK) {
    cudaError_t e2 = cudaMemcpyAsync(st->d, scratch, bytes, cudaMemcpyDeviceToDevice, ctx->stream);
    if (e2 == cudaSuccess) e2 = cudaStreamSynchronize(ctx->stream);
    if (e2 != cudaSuccess) {
      iqsb_set_error("iqsb_permute_local: copy back failed: %s", cudaGetErrorString(e2));
      rc = IQSB_ERR_CUDA;
    }
  }
  cudaFree(scratch);
  return rc;
}
