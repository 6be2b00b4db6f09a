// kernels_reduce.cu -- reductions over the state vector (warp shuffle -> block -> fixed-order
// second stage, so results are deterministic run to run).
//
// Reference loops replaced:
//   GetProbability      src/qureg_measure.cpp:150-167   (serial sum on the CPU)
//   ExpectationValue    src/qureg_expectval.cpp:173-185 (signed |a|^2 by parity of popcount)
//   ComputeNorm         src/qureg_utils.cpp:241-245
//   ComputeOverlap      src/qureg_utils.cpp:282-287
//   MaxAbsDiff          src/qureg_utils.cpp:54-57
//   MaxL2NormDiff       src/qureg_utils.cpp:142-146
//   IsClassicalBit / GetClassicalValue  src/qureg_measure.cpp:34-78, 198-240
//   operator==          src/qureg_utils.cpp:21-29
//   Entropy / GoogleStats  src/qureg_utils.cpp:316-324, 372-403
// Accumulation is always in double, also for float registers.
#include "iqsb_internal.cuh"

namespace {

constexpr int kBlock = 256;
constexpr int kWarps = kBlock / 32;

template <int NOUT, bool IS_MAX>
__device__ __forceinline__ void block_reduce_store(double (&acc)[NOUT], double *partials) {
  __shared__ double sm[NOUT][kWarps];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NOUT; ++k) {
    double v = acc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      double w = __shfl_xor_sync(0xffffffffu, v, o);
      v = IS_MAX ? fmax(v, w) : v + w;
    }
    if (lane == 0) sm[k][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < NOUT) {
    double v = sm[threadIdx.x][0];
    for (int w = 1; w < kWarps; ++w) v = IS_MAX ? fmax(v, sm[threadIdx.x][w]) : v + sm[threadIdx.x][w];
    partials[(size_t)blockIdx.x * NOUT + threadIdx.x] = v;
  }
}

// second stage: one block, fixed order over the per-block partials
template <int NOUT, bool IS_MAX>
__global__ void k_finish(const double *partials, int nblocks, double *result) {
  double acc[NOUT];
#pragma unroll
  for (int k = 0; k < NOUT; ++k) acc[k] = IS_MAX ? -1.0 : 0.0;
  for (int b = threadIdx.x; b < nblocks; b += kBlock)
#pragma unroll
    for (int k = 0; k < NOUT; ++k) {
      double v = partials[(size_t)b * NOUT + k];
      acc[k] = IS_MAX ? fmax(acc[k], v) : acc[k] + v;
    }
  __shared__ double sm[NOUT][kWarps];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NOUT; ++k) {
    double v = acc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      double w = __shfl_xor_sync(0xffffffffu, v, o);
      v = IS_MAX ? fmax(v, w) : v + w;
    }
    if (lane == 0) sm[k][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < NOUT) {
    double v = sm[threadIdx.x][0];
    for (int w = 1; w < kWarps; ++w) v = IS_MAX ? fmax(v, sm[threadIdx.x][w]) : v + sm[threadIdx.x][w];
    result[threadIdx.x] = v;
  }
}

template <typename T>
__device__ __forceinline__ double norm_d(Cx<T> a) {
  double re = (double)a.re, im = (double)a.im;
  return __dadd_rn(__dmul_rn(re, re), __dmul_rn(im, im));
}

// sum |a|^2 over a subset (Geom in chunk units, off0 = fixed bits)
template <typename T>
__global__ void __launch_bounds__(kBlock) k_norm_subset_w2(const Chunk<T> *__restrict__ s, Geom g, double *partials) {
  double acc[1] = {0.0};
  const uint64_t stride = (uint64_t)gridDim.x * kBlock;
  uint64_t t = (uint64_t)blockIdx.x * kBlock + threadIdx.x;
  for (; t + 3 * stride < g.nwork; t += 4 * stride) {
    Chunk<T> v0 = ld_chunk(s + expand(t, g) + g.off0);
    Chunk<T> v1 = ld_chunk(s + expand(t + stride, g) + g.off0);
    Chunk<T> v2 = ld_chunk(s + expand(t + 2 * stride, g) + g.off0);
    Chunk<T> v3 = ld_chunk(s + expand(t + 3 * stride, g) + g.off0);
    acc[0] += (norm_d(v0.a) + norm_d(v0.b)) + (norm_d(v1.a) + norm_d(v1.b));
    acc[0] += (norm_d(v2.a) + norm_d(v2.b)) + (norm_d(v3.a) + norm_d(v3.b));
  }
  for (; t < g.nwork; t += stride) {
    Chunk<T> v = ld_chunk(s + expand(t, g) + g.off0);
    acc[0] += norm_d(v.a) + norm_d(v.b);
  }
  block_reduce_store<1, false>(acc, partials);
}
template <typename T>
__global__ void __launch_bounds__(kBlock) k_norm_subset_w1(const Cx<T> *__restrict__ s, Geom g, double *partials) {
  double acc[1] = {0.0};
  const uint64_t stride = (uint64_t)gridDim.x * kBlock;
  for (uint64_t t = (uint64_t)blockIdx.x * kBlock + threadIdx.x; t < g.nwork; t += stride)
    acc[0] += norm_d(ld_amp(s + expand(t, g) + g.off0));
  block_reduce_store<1, false>(acc, partials);
}

// sum (-1)^popcount((glb_start + i) & mask) |a_i|^2
template <typename T>
__global__ void __launch_bounds__(kBlock)
    k_parity(const Chunk<T> *__restrict__ s, uint64_t nchunks, uint64_t mask, uint64_t glb_start, double *partials) {
  double acc[1] = {0.0};
  const uint64_t stride = (uint64_t)gridDim.x * kBlock;
  uint64_t t = (uint64_t)blockIdx.x * kBlock + threadIdx.x;
  auto term = [&](uint64_t c, const Chunk<T> &v) {
    uint64_t ia = glb_start + 2 * c;
    double na = norm_d(v.a), nb = norm_d(v.b);
    double sa = (__popcll(ia & mask) & 1) ? -na : na;
    double sb = (__popcll((ia + 1) & mask) & 1) ? -nb : nb;
    return sa + sb;
  };
  for (; t + 3 * stride < nchunks; t += 4 * stride) {
    Chunk<T> v0 = ld_chunk(s + t), v1 = ld_chunk(s + t + stride);
    Chunk<T> v2 = ld_chunk(s + t + 2 * stride), v3 = ld_chunk(s + t + 3 * stride);
    acc[0] += (term(t, v0) + term(t + stride, v1)) + (term(t + 2 * stride, v2) + term(t + 3 * stride, v3));
  }
  for (; t < nchunks; t += stride) acc[0] += term(t, ld_chunk(s + t));
  block_reduce_store<1, false>(acc, partials);
}

// two-register reductions. MODE 0: overlap sum conj(b) a -> (re, im); 1: max |a - f b|; 2: sum |a-b|^2
template <typename T, int MODE>
__global__ void __launch_bounds__(kBlock)
    k_two(const Chunk<T> *__restrict__ a, const Chunk<T> *__restrict__ b, uint64_t nchunks, Cx<T> f, double *partials) {
  constexpr int NOUT = MODE == 0 ? 2 : 1;
  double acc[NOUT];
#pragma unroll
  for (int k = 0; k < NOUT; ++k) acc[k] = MODE == 1 ? -1.0 : 0.0;
  const uint64_t stride = (uint64_t)gridDim.x * kBlock;
  auto one = [&](Cx<T> x, Cx<T> y) {
    if (MODE == 0) {
      // conj(psi) * state, psi = y (reference qureg_utils.cpp:284): (c - id)(a + ib)
      Cx<T> yc = {y.re, -y.im};
      Cx<T> p = cmul(yc, x);
      acc[0] += (double)p.re;
      acc[1] += (double)p.im;
    } else if (MODE == 1) {
      Cx<T> fy = cmul(f, y);
      double dr = (double)sub_rn(x.re, fy.re), di = (double)sub_rn(x.im, fy.im);
      acc[0] = fmax(acc[0], hypot(dr, di));
    } else {
      T dr = sub_rn(x.re, y.re), di = sub_rn(x.im, y.im);
      acc[0] += norm_d(Cx<T>{dr, di});
    }
  };
  uint64_t t = (uint64_t)blockIdx.x * kBlock + threadIdx.x;
  for (; t + stride < nchunks; t += 2 * stride) {
    Chunk<T> x0 = ld_chunk(a + t), y0 = ld_chunk(b + t);
    Chunk<T> x1 = ld_chunk(a + t + stride), y1 = ld_chunk(b + t + stride);
    one(x0.a, y0.a); one(x0.b, y0.b); one(x1.a, y1.a); one(x1.b, y1.b);
  }
  for (; t < nchunks; t += stride) {
    Chunk<T> x0 = ld_chunk(a + t), y0 = ld_chunk(b + t);
    one(x0.a, y0.a); one(x0.b, y0.b);
  }
  block_reduce_store<NOUT, MODE == 1>(acc, partials);
}

// flags[0] |= any |a|^2 > tol with selected bit 0 ; flags[1] likewise with bit 1.
template <typename T>
__global__ void __launch_bounds__(kBlock)
    k_any_above(const Cx<T> *__restrict__ s, uint64_t n, unsigned pos, unsigned cbit, double tol, int *flags) {
  int f0 = 0, f1 = 0;
  const uint64_t stride = (uint64_t)gridDim.x * kBlock;
  for (uint64_t i = (uint64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += stride) {
    Cx<T> v = ld_amp(s + i);
    // std::norm in the register's own precision, as the reference does
    bool big = (double)cnorm(v) > tol;
    unsigned bit = pos < 64 ? (unsigned)((i >> pos) & 1) : cbit;
    if (big) { if (bit) f1 = 1; else f0 = 1; }
  }
  f0 = __any_sync(0xffffffffu, f0);
  f1 = __any_sync(0xffffffffu, f1);
  if ((threadIdx.x & 31) == 0) {
    if (f0) atomicOr(flags + 0, 1);
    if (f1) atomicOr(flags + 1, 1);
  }
}

template <typename T>
__global__ void __launch_bounds__(kBlock)
    k_not_equal(const Chunk<T> *__restrict__ a, const Chunk<T> *__restrict__ b, uint64_t nchunks, int *flags) {
  int ne = 0;
  const uint64_t stride = (uint64_t)gridDim.x * kBlock;
  for (uint64_t t = (uint64_t)blockIdx.x * kBlock + threadIdx.x; t < nchunks; t += stride) {
    Chunk<T> x = ld_chunk(a + t), y = ld_chunk(b + t);
    if (x.a.re != y.a.re || x.a.im != y.a.im || x.b.re != y.b.re || x.b.im != y.b.im) ne = 1;
  }
  ne = __any_sync(0xffffffffu, ne);
  if (ne && (threadIdx.x & 31) == 0) atomicOr(flags + 0, 1);
}

// out[0] = sum -p ln p, out[1] = sum -ln p (p != 0), out[2..10] = sum p^k, k=2..10
template <typename T>
__global__ void __launch_bounds__(kBlock) k_entropy(const Cx<T> *__restrict__ s, uint64_t n, double *partials) {
  double acc[11];
#pragma unroll
  for (int k = 0; k < 11; ++k) acc[k] = 0.0;
  const uint64_t stride = (uint64_t)gridDim.x * kBlock;
  for (uint64_t i = (uint64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += stride) {
    double pj = (double)cnorm(ld_amp(s + i));
    if (pj != 0.0) {
      double nl = log(pj);
      acc[0] -= pj * nl;
      acc[1] -= nl;
    }
    double pj2 = pj * pj, pj3 = pj2 * pj, pj4 = pj2 * pj2, pj5 = pj3 * pj2, pj6 = pj3 * pj3,
           pj7 = pj4 * pj3, pj8 = pj4 * pj4, pj9 = pj5 * pj4, pj10 = pj5 * pj5;
    acc[2] += pj2; acc[3] += pj3; acc[4] += pj4; acc[5] += pj5; acc[6] += pj6;
    acc[7] += pj7; acc[8] += pj8; acc[9] += pj9; acc[10] += pj10;
  }
  block_reduce_store<11, false>(acc, partials);
}

// All marginals in one sweep (16 B per amplitude read once): out[0] = sum |a|^2, out[1 + q] = sum of
// |a_i|^2 over the local indices i with bit q set.  The grid has a power-of-two number of threads,
// thread t visits chunks t, t + 2^S, t + 2 * 2^S, ...: the low S chunk bits are the thread's own
// (credited once, at the end, from its total), bit 0 of the amplitude index is the position inside the
// chunk, and only the few bits above S cost a predicated add per chunk.
constexpr int kProbOut = 40;   // 1 + up to 39 local positions
constexpr int kProbHigh = 16;  // positions above the thread-index bits
constexpr int kProbLog2Threads = 18;  // full grid: 2^18 threads = 1024 blocks, all resident
template <typename T>
__global__ void __launch_bounds__(kBlock)
    k_prob_all(const Chunk<T> *__restrict__ s, uint64_t nchunks, unsigned log2_threads, unsigned nbits, double *partials) {
  const uint64_t t = (uint64_t)blockIdx.x * kBlock + threadIdx.x;
  const uint64_t stride = 1ull << log2_threads;
  double total = 0.0, odd = 0.0, hi[kProbHigh];
#pragma unroll
  for (int j = 0; j < kProbHigh; ++j) hi[j] = 0.0;
  auto add = [&](uint64_t k, const Chunk<T> &v) {
    const double na = norm_d(v.a), nb = norm_d(v.b), p = na + nb;
    total += p;
    odd += nb;
#pragma unroll
    for (int j = 0; j < kProbHigh; ++j)
      if ((k >> j) & 1ull) hi[j] += p;
  };
  if (t < nchunks) {
    const uint64_t niter = nchunks >> log2_threads;
    uint64_t k = 0;
    for (; k + 3 < niter; k += 4) {
      Chunk<T> v0 = ld_chunk(s + t + k * stride), v1 = ld_chunk(s + t + (k + 1) * stride);
      Chunk<T> v2 = ld_chunk(s + t + (k + 2) * stride), v3 = ld_chunk(s + t + (k + 3) * stride);
      add(k, v0); add(k + 1, v1); add(k + 2, v2); add(k + 3, v3);
    }
    for (; k < niter; ++k) add(k, ld_chunk(s + t + k * stride));
  }
  double acc[kProbOut];
#pragma unroll
  for (int q = 0; q < kProbOut; ++q) acc[q] = 0.0;
  acc[0] = total;
  acc[1] = odd;
#pragma unroll
  for (int q = 1; q < kProbOut - 1; ++q) {  // amplitude bit q = chunk bit q - 1
    const unsigned cb = (unsigned)q - 1u;
    double v = 0.0;
    if ((unsigned)q < nbits) {
      // bits above the thread index exist only with the full grid (log2_threads == kProbLog2Threads)
      if (cb < log2_threads) v = ((t >> cb) & 1ull) ? total : 0.0;
      else if (cb >= (unsigned)kProbLog2Threads && cb - kProbLog2Threads < (unsigned)kProbHigh) v = hi[cb >= (unsigned)kProbLog2Threads ? cb - kProbLog2Threads : 0];
    }
    acc[1 + q] = v;
  }
  block_reduce_store<kProbOut, false>(acc, partials);
}

// Read-only expectation value of a Pauli string (X on the bits of xmask, Y on ymask, Z on zmask):
// P|i> = i^ny (-1)^popc(i & (y|z)) |i ^ f>, f = x | y, so <psi|P|psi> = i^ny * S with
// S = sum_i s_i psi_i conj(psi_{i^f}).  The pair (i, i^f) contributes s_i (w + (-1)^ny conj(w)),
// w = psi_i conj(psi_{i^f}); the kernel visits every pair once (16 B per amplitude, nothing written)
// and accumulates sum s_i Re w and sum s_i Im w.  Work item = chunk c with the highest flip bit clear,
// partner chunk c ^ (f >> 1), its two amplitudes swapped when f has bit 0.
template <typename T, bool INCHUNK>
__global__ void __launch_bounds__(kBlock)
    k_pauli(const Chunk<T> *__restrict__ s, uint64_t nwork, unsigned hb, uint64_t fchunk, int swap, uint64_t smask, uint64_t glb_start, double *partials) {
  double acc[3] = {0.0, 0.0, 0.0};  // sum s Re w, sum s Im w, sum |psi|^2 (every amplitude is read exactly once)
  const uint64_t stride = (uint64_t)gridDim.x * kBlock;
  auto pair = [&](uint64_t i, Cx<T> a, Cx<T> b) {  // w = a conj(b), sign from the global index of a
    const double ar = (double)a.re, ai = (double)a.im, br = (double)b.re, bi = (double)b.im;
    const double wr = ar * br + ai * bi, wi = ai * br - ar * bi;
    const bool neg = __popcll((glb_start + i) & smask) & 1;
    acc[0] += neg ? -wr : wr;
    acc[1] += neg ? -wi : wi;
    acc[2] += (ar * ar + ai * ai) + (br * br + bi * bi);
  };
  auto one = [&](uint64_t t2, const Chunk<T> &A, const Chunk<T> &B) {
    const uint64_t c = INCHUNK ? t2 : insert_zero(t2, hb);
    if (INCHUNK) {
      pair(2 * c, A.a, A.b);
    } else {
      pair(2 * c, A.a, swap ? B.b : B.a);
      pair(2 * c + 1, A.b, swap ? B.a : B.b);
    }
  };
  uint64_t t = (uint64_t)blockIdx.x * kBlock + threadIdx.x;
  for (; t + stride < nwork; t += 2 * stride) {
    const uint64_t c0 = INCHUNK ? t : insert_zero(t, hb), c1 = INCHUNK ? t + stride : insert_zero(t + stride, hb);
    Chunk<T> A0 = ld_chunk(s + c0), A1 = ld_chunk(s + c1), B0 = A0, B1 = A1;
    if (!INCHUNK) {
      B0 = ld_chunk(s + (c0 ^ fchunk));
      B1 = ld_chunk(s + (c1 ^ fchunk));
    }
    one(t, A0, B0);
    one(t + stride, A1, B1);
  }
  for (; t < nwork; t += stride) {
    const uint64_t c0 = INCHUNK ? t : insert_zero(t, hb);
    Chunk<T> A0 = ld_chunk(s + c0), B0 = A0;
    if (!INCHUNK) B0 = ld_chunk(s + (c0 ^ fchunk));
    one(t, A0, B0);
  }
  block_reduce_store<3, false>(acc, partials);
}

inline int red_grid(const iqsb_ctx *ctx, uint64_t nwork) {
  uint64_t want = (nwork + kBlock - 1) / kBlock;
  int cap = ctx->num_sms * 8;
  if (cap > kMaxRedBlocks) cap = kMaxRedBlocks;
  if (want < 1) want = 1;
  return (int)(want < (uint64_t)cap ? want : (uint64_t)cap);
}

template <int NOUT, bool IS_MAX>
int finish(iqsb_ctx *ctx, int nblocks, double *out) {
  k_finish<NOUT, IS_MAX><<<1, kBlock, 0, ctx->stream>>>(ctx->d_partials, nblocks, ctx->d_result);
  IQSB_TRY(iqsb_check_launch(ctx, "k_finish"));
  IQSB_CUDA(cudaMemcpyAsync(ctx->h_result, ctx->d_result, NOUT * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  IQSB_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int k = 0; k < NOUT; ++k) out[k] = ctx->h_result[k];
  return iqsb_check(ctx);
}

int fetch_flags(iqsb_ctx *ctx, int n, int *out) {
  IQSB_CUDA(cudaMemcpyAsync(ctx->h_result, ctx->d_flags, n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  IQSB_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int k = 0; k < n; ++k) out[k] = ((int *)ctx->h_result)[k];
  return IQSB_OK;
}

}  // namespace

static int norm_subset(iqsb_state *st, int nfix, const unsigned *pos_sorted, const unsigned *val, double *out) {
  iqsb_ctx *ctx = st->ctx;
  bool w1 = (nfix > 0 && pos_sorted[0] == 0) || st->local_amps < 2;
  unsigned ins[3];
  uint64_t off = 0;
  int grid;
  if (w1) {
    for (int i = 0; i < nfix; ++i) { ins[i] = pos_sorted[i]; off |= (uint64_t)val[i] << pos_sorted[i]; }
    Geom g = make_geom(st->local_amps >> nfix, nfix, ins, off, 0);
    grid = red_grid(ctx, g.nwork);
    if (st->dtype == IQSB_F64)
      k_norm_subset_w1<double><<<grid, kBlock, 0, ctx->stream>>>((const Cx<double> *)st->d, g, ctx->d_partials);
    else
      k_norm_subset_w1<float><<<grid, kBlock, 0, ctx->stream>>>((const Cx<float> *)st->d, g, ctx->d_partials);
  } else {
    for (int i = 0; i < nfix; ++i) { ins[i] = pos_sorted[i] - 1; off |= (uint64_t)val[i] << (pos_sorted[i] - 1); }
    Geom g = make_geom((st->local_amps / 2) >> nfix, nfix, ins, off, 0);
    grid = red_grid(ctx, g.nwork);
    if (st->dtype == IQSB_F64)
      k_norm_subset_w2<double><<<grid, kBlock, 0, ctx->stream>>>((const Chunk<double> *)st->d, g, ctx->d_partials);
    else
      k_norm_subset_w2<float><<<grid, kBlock, 0, ctx->stream>>>((const Chunk<float> *)st->d, g, ctx->d_partials);
  }
  IQSB_TRY(iqsb_check_launch(ctx, "k_norm_subset"));
  return finish<1, false>(ctx, grid, out);
}

extern "C" int iqsb_norm2(iqsb_state *st, double *out) {
  IQSB_REQUIRE(st && out, "iqsb_norm2: null argument");
  return norm_subset(st, 0, nullptr, nullptr, out);
}

extern "C" int iqsb_prob1(iqsb_state *st, unsigned pos, double *out) {
  IQSB_REQUIRE(st && out, "iqsb_prob1: null argument");
  IQSB_REQUIRE(pos < st->log2_local, "iqsb_prob1: position must be local");
  unsigned p[1] = {pos}, v[1] = {1};
  return norm_subset(st, 1, p, v, out);
}

extern "C" int iqsb_parity_expect(iqsb_state *st, uint64_t mask, uint64_t glb_start, double *out) {
  IQSB_REQUIRE(st && out, "iqsb_parity_expect: null argument");
  IQSB_REQUIRE(st->local_amps >= 2, "iqsb_parity_expect: shard too small");
  iqsb_ctx *ctx = st->ctx;
  uint64_t nchunks = st->local_amps / 2;
  int grid = red_grid(ctx, nchunks);
  if (st->dtype == IQSB_F64)
    k_parity<double><<<grid, kBlock, 0, ctx->stream>>>((const Chunk<double> *)st->d, nchunks, mask, glb_start, ctx->d_partials);
  else
    k_parity<float><<<grid, kBlock, 0, ctx->stream>>>((const Chunk<float> *)st->d, nchunks, mask, glb_start, ctx->d_partials);
  IQSB_TRY(iqsb_check_launch(ctx, "k_parity"));
  return finish<1, false>(ctx, grid, out);
}

extern "C" int iqsb_prob_all(iqsb_state *st, double *out, int nout) {
  IQSB_REQUIRE(st && out, "iqsb_prob_all: null argument");
  const unsigned M = st->log2_local;
  IQSB_REQUIRE(nout >= (int)M + 1, "iqsb_prob_all: out[] must hold 1 + log2(local_amps) = %u doubles", M + 1);
  IQSB_REQUIRE(M >= 1 && M + 1 <= (unsigned)kProbOut - 1, "iqsb_prob_all: shard of 2^%u amplitudes is not supported", M);
  iqsb_ctx *ctx = st->ctx;
  const uint64_t nchunks = st->local_amps / 2;
  // a power-of-two number of threads, at most 1024 blocks (all resident: 148 SMs x 8)
  unsigned log2_threads = kProbLog2Threads;
  if (M - 1 < log2_threads) log2_threads = M - 1;
  IQSB_REQUIRE(M - 1 - log2_threads <= (unsigned)kProbHigh, "iqsb_prob_all: shard too large");
  const uint64_t nthreads = 1ull << log2_threads;
  const int grid = (int)((nthreads + kBlock - 1) / kBlock);
  if (st->dtype == IQSB_F64)
    k_prob_all<double><<<grid, kBlock, 0, ctx->stream>>>((const Chunk<double> *)st->d, nchunks, log2_threads, M, ctx->d_partials);
  else
    k_prob_all<float><<<grid, kBlock, 0, ctx->stream>>>((const Chunk<float> *)st->d, nchunks, log2_threads, M, ctx->d_partials);
  IQSB_TRY(iqsb_check_launch(ctx, "k_prob_all", (double)st->local_amps * st->amp_bytes()));
  double r[kProbOut];
  IQSB_TRY((finish<kProbOut, false>(ctx, grid, r)));
  for (unsigned q = 0; q <= M; ++q) out[q] = r[q];
  return IQSB_OK;
}

extern "C" int iqsb_pauli_expect(iqsb_state *st, uint64_t xmask, uint64_t ymask, uint64_t zmask, uint64_t glb_start, double out[2]) {
  IQSB_REQUIRE(st && out, "iqsb_pauli_expect: null argument");
  IQSB_REQUIRE(!(xmask & ymask) && !(xmask & zmask) && !(ymask & zmask), "iqsb_pauli_expect: a position carries two observables");
  IQSB_REQUIRE(st->local_amps >= 2, "iqsb_pauli_expect: shard too small");
  const uint64_t f = xmask | ymask;
  IQSB_REQUIRE(f < st->local_amps, "iqsb_pauli_expect: X / Y observables must sit on local positions");
  if (f == 0) {
    IQSB_TRY(iqsb_parity_expect(st, zmask, glb_start, &out[0]));
    return iqsb_norm2(st, &out[1]);
  }
  iqsb_ctx *ctx = st->ctx;
  const uint64_t smask = ymask | zmask;
  const int ny = __builtin_popcountll(ymask);
  const uint64_t fchunk = f >> 1;
  const int swap = (int)(f & 1);
  double r[3] = {0., 0., 0.};
  int grid;
  if (fchunk == 0) {  // X or Y on position 0 only: both partners in one chunk
    const uint64_t nwork = st->local_amps / 2;
    grid = red_grid(ctx, nwork);
    if (st->dtype == IQSB_F64)
      k_pauli<double, true><<<grid, kBlock, 0, ctx->stream>>>((const Chunk<double> *)st->d, nwork, 0, 0, 0, smask, glb_start, ctx->d_partials);
    else
      k_pauli<float, true><<<grid, kBlock, 0, ctx->stream>>>((const Chunk<float> *)st->d, nwork, 0, 0, 0, smask, glb_start, ctx->d_partials);
  } else {
    const unsigned hb = 63u - (unsigned)__builtin_clzll(fchunk);
    const uint64_t nwork = st->local_amps / 4;
    grid = red_grid(ctx, nwork);
    if (st->dtype == IQSB_F64)
      k_pauli<double, false><<<grid, kBlock, 0, ctx->stream>>>((const Chunk<double> *)st->d, nwork, hb, fchunk, swap, smask, glb_start, ctx->d_partials);
    else
      k_pauli<float, false><<<grid, kBlock, 0, ctx->stream>>>((const Chunk<float> *)st->d, nwork, hb, fchunk, swap, smask, glb_start, ctx->d_partials);
  }
  IQSB_TRY(iqsb_check_launch(ctx, "k_pauli", (double)st->local_amps * st->amp_bytes()));
  IQSB_TRY((finish<3, false>(ctx, grid, r)));
  // <P> = Re(i^ny S); S = 2 sum s Re w (ny even) or 2 i sum s Im w (ny odd)
  switch (ny & 3) {
    case 0: out[0] = 2.0 * r[0]; break;
    case 1: out[0] = -2.0 * r[1]; break;
    case 2: out[0] = -2.0 * r[0]; break;
    default: out[0] = 2.0 * r[1]; break;
  }
  out[1] = r[2];
  return IQSB_OK;
}

template <int MODE>
static int two_reg(iqsb_state *a, iqsb_state *b, const double f[2], double *out) {
  IQSB_REQUIRE(a && b && out, "two-register reduction: null argument");
  IQSB_REQUIRE(a->local_amps == b->local_amps && a->dtype == b->dtype && a->ctx == b->ctx,
               "two-register reduction: registers do not match");
  if (a->local_amps == 1) {  // the reference's default-constructed register: one amplitude, on the host
    double xr, xi, yr, yi;
    IQSB_TRY(iqsb_get_amp(a, 0, &xr, &xi));
    IQSB_TRY(iqsb_get_amp(b, 0, &yr, &yi));
    if (MODE == 0) {  // conj(b) a
      out[0] = yr * xr + yi * xi;
      out[1] = yr * xi - yi * xr;
    } else if (MODE == 1) {
      const double fr = f[0] * yr - f[1] * yi, fi = f[0] * yi + f[1] * yr;
      out[0] = hypot(xr - fr, xi - fi);
    } else {
      out[0] = (xr - yr) * (xr - yr) + (xi - yi) * (xi - yi);
    }
    return IQSB_OK;
  }
  iqsb_ctx *ctx = a->ctx;
  uint64_t nchunks = a->local_amps / 2;
  int grid = red_grid(ctx, nchunks);
  if (a->dtype == IQSB_F64)
    k_two<double, MODE><<<grid, kBlock, 0, ctx->stream>>>((const Chunk<double> *)a->d, (const Chunk<double> *)b->d,
                                                          nchunks, Cx<double>{f[0], f[1]}, ctx->d_partials);
  else
    k_two<float, MODE><<<grid, kBlock, 0, ctx->stream>>>((const Chunk<float> *)a->d, (const Chunk<float> *)b->d,
                                                         nchunks, Cx<float>{(float)f[0], (float)f[1]}, ctx->d_partials);
  IQSB_TRY(iqsb_check_launch(ctx, "k_two"));
  if (MODE == 0) return finish<2, false>(ctx, grid, out);
  if (MODE == 1) return finish<1, true>(ctx, grid, out);
  return finish<1, false>(ctx, grid, out);
}

extern "C" int iqsb_overlap(iqsb_state *a, iqsb_state *b, double out[2]) {
  const double one[2] = {1, 0};
  return two_reg<0>(a, b, one, out);
}
extern "C" int iqsb_maxabsdiff(iqsb_state *a, iqsb_state *b, const double s[2], double *out) {
  IQSB_REQUIRE(s, "iqsb_maxabsdiff: null factor");
  return two_reg<1>(a, b, s, out);
}
extern "C" int iqsb_l2diff(iqsb_state *a, iqsb_state *b, double *out) {
  const double one[2] = {1, 0};
  return two_reg<2>(a, b, one, out);
}

extern "C" int iqsb_any_above(iqsb_state *st, unsigned pos, double tol, uint64_t glb_start, int out[2]) {
  IQSB_REQUIRE(st && out, "iqsb_any_above: null argument");
  iqsb_ctx *ctx = st->ctx;
  IQSB_CUDA(cudaMemsetAsync(ctx->d_flags, 0, 4 * sizeof(int), ctx->stream));
  unsigned p = pos < st->log2_local ? pos : 64;
  unsigned cbit = pos < st->log2_local ? 0 : (unsigned)((glb_start >> pos) & 1);
  int grid = red_grid(ctx, st->local_amps);
  if (st->dtype == IQSB_F64)
    k_any_above<double><<<grid, kBlock, 0, ctx->stream>>>((const Cx<double> *)st->d, st->local_amps, p, cbit, tol, ctx->d_flags);
  else
    k_any_above<float><<<grid, kBlock, 0, ctx->stream>>>((const Cx<float> *)st->d, st->local_amps, p, cbit, tol, ctx->d_flags);
  IQSB_TRY(iqsb_check_launch(ctx, "k_any_above"));
  return fetch_flags(ctx, 2, out);
}

extern "C" int iqsb_equal(iqsb_state *a, iqsb_state *b, int *out) {
  IQSB_REQUIRE(a && b && out, "iqsb_equal: null argument");
  IQSB_REQUIRE(a->local_amps == b->local_amps && a->dtype == b->dtype && a->ctx == b->ctx,
               "iqsb_equal: registers do not match");
  if (a->local_amps == 1) {
    double xr, xi, yr, yi;
    IQSB_TRY(iqsb_get_amp(a, 0, &xr, &xi));
    IQSB_TRY(iqsb_get_amp(b, 0, &yr, &yi));
    *out = (xr == yr && xi == yi) ? 1 : 0;
    return IQSB_OK;
  }
  iqsb_ctx *ctx = a->ctx;
  IQSB_CUDA(cudaMemsetAsync(ctx->d_flags, 0, 4 * sizeof(int), ctx->stream));
  uint64_t nchunks = a->local_amps / 2;
  int grid = red_grid(ctx, nchunks);
  if (a->dtype == IQSB_F64)
    k_not_equal<double><<<grid, kBlock, 0, ctx->stream>>>((const Chunk<double> *)a->d, (const Chunk<double> *)b->d, nchunks, ctx->d_flags);
  else
    k_not_equal<float><<<grid, kBlock, 0, ctx->stream>>>((const Chunk<float> *)a->d, (const Chunk<float> *)b->d, nchunks, ctx->d_flags);
  IQSB_TRY(iqsb_check_launch(ctx, "k_not_equal"));
  int ne = 0;
  IQSB_TRY(fetch_flags(ctx, 1, &ne));
  *out = ne ? 0 : 1;
  return IQSB_OK;
}

extern "C" int iqsb_entropy_stats(iqsb_state *st, double out[11]) {
  IQSB_REQUIRE(st && out, "iqsb_entropy_stats: null argument");
  iqsb_ctx *ctx = st->ctx;
  int grid = red_grid(ctx, st->local_amps);
  if (st->dtype == IQSB_F64)
    k_entropy<double><<<grid, kBlock, 0, ctx->stream>>>((const Cx<double> *)st->d, st->local_amps, ctx->d_partials);
  else
    k_entropy<float><<<grid, kBlock, 0, ctx->stream>>>((const Cx<float> *)st->d, st->local_amps, ctx->d_partials);
  IQSB_TRY(iqsb_check_launch(ctx, "k_entropy"));
  return finish<11, false>(ctx, grid, out);
}
