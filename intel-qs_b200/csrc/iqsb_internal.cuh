// iqsb_internal.cuh -- shared definitions of the sm_100a state-vector engine.
//
// Data layout in HBM: one shard = `local_amps` interleaved complex numbers (re, im) of
// T = double (16 B per amplitude) or float (8 B).  The basic unit of memory traffic is a
// "chunk" of TWO adjacent amplitudes: 32 B for double, moved with one 256-bit
// LDG.E.ENL2.256 / STG.E.ENL2.256 (sm_100 and newer), 16 B for float (LDG.128).
// A warp therefore touches 1 KiB (double) of contiguous memory per instruction.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/iqsb.h"

// ---------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------
void iqsb_set_error(const char *fmt, ...);

#define IQSB_CUDA(call)                                                                     \
  do {                                                                                      \
    cudaError_t e__ = (call);                                                               \
    if (e__ != cudaSuccess) {                                                               \
      iqsb_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return IQSB_ERR_CUDA;                                                                 \
    }                                                                                       \
  } while (0)

#define IQSB_REQUIRE(cond, ...)         \
  do {                                  \
    if (!(cond)) {                      \
      iqsb_set_error(__VA_ARGS__);      \
      return IQSB_ERR_ARG;              \
    }                                   \
  } while (0)

#define IQSB_TRY(call)            \
  do {                            \
    int r__ = (call);             \
    if (r__ != IQSB_OK) return r__; \
  } while (0)

// ---------------------------------------------------------------------------------------
// host-side objects behind the opaque handles
// ---------------------------------------------------------------------------------------
struct iqsb_peer_table;  // comm.cu
struct iqsb_prof;        // api.cu: per-kernel-class device timing (iqsb_profile)

struct iqsb_ctx {
  int rank = 0, nranks = 1, device = 0, num_sms = 148;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;  // stream in use (own or adopted)
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t *slots = nullptr;  // numbered events (iqsb_event_record)
  uint64_t launches = 0;
  uint64_t nvlink_bytes = 0;
  // reduction scratch: per-block partials + final results (device), pinned host mirror
  double *d_partials = nullptr;  // [kMaxRedBlocks * kMaxRedOut]
  double *d_result = nullptr;    // [kMaxRedOut]
  double *h_result = nullptr;    // pinned
  int *d_flags = nullptr;        // [4]
  void *comm = nullptr;          // ncclComm_t
  iqsb_peer_table *peers = nullptr;
  int arith = IQSB_ARITH_EXACT;  // iqsb_set_arith
  iqsb_prof *prof = nullptr;     // iqsb_profile
  unsigned long long *d_tile_counter = nullptr;  // tile scheduler of the fused kernel
  // status word written by kernels that give up (the peer barrier's deadline): pinned host memory
  // mapped into the device, read by iqsb_check
  int *h_status = nullptr, *d_status = nullptr;
  double barrier_timeout_s = 300.;  // IQS_B200_BARRIER_TIMEOUT_S; 0 = wait for ever
};

struct iqsb_state {
  iqsb_ctx *ctx = nullptr;
  void *d = nullptr;  // device (or managed) pointer to the shard, followed by the tmp area
  uint64_t local_amps = 0, tmp_amps = 0;
  int dtype = IQSB_F64, mem_kind = IQSB_MEM_DEVICE;
  unsigned log2_local = 0;
  // distributed: peer-mapped pointers to the same register's shard on every rank
  void **peer_ptr = nullptr;  // [nranks], peer_ptr[rank] == d
  bool shared = false;
  size_t amp_bytes() const { return dtype == IQSB_F64 ? 16 : 8; }
};

constexpr int kMaxRedBlocks = 148 * 8;
constexpr int kMaxRedOut = 40;
constexpr int kMaxEventSlots = 4096;

// ---------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------
#ifdef __CUDACC__

// Arithmetic is deliberately NOT contracted into FMAs: the reference is compiled for
// baseline x86-64 (no FMA) and evaluates (a+ib)(c+id) = (ac - bd) + i(ad + bc) with
// separately rounded products (libstdc++ std::complex, reference SURVEY 8c).  Using the
// _rn intrinsics keeps the GPU result bit-identical to that evaluation order.  The kernels
// are HBM-bound; the extra FP64 instructions are free.
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }

template <typename T>
struct Cx {
  T re, im;
};

template <typename T>
__device__ __forceinline__ Cx<T> cmul(Cx<T> a, Cx<T> b) {
  Cx<T> r;
  r.re = sub_rn(mul_rn(a.re, b.re), mul_rn(a.im, b.im));
  r.im = add_rn(mul_rn(a.re, b.im), mul_rn(a.im, b.re));
  return r;
}
template <typename T>
__device__ __forceinline__ Cx<T> cadd(Cx<T> a, Cx<T> b) {
  Cx<T> r;
  r.re = add_rn(a.re, b.re);
  r.im = add_rn(a.im, b.im);
  return r;
}
template <typename T>
__device__ __forceinline__ T cnorm(Cx<T> a) {
  return add_rn(mul_rn(a.re, a.re), mul_rn(a.im, a.im));
}

// 2x2 complex matrix passed by value as a kernel argument
template <typename T>
struct Mat2 {
  Cx<T> m00, m01, m10, m11;
};

// out0 = m00*in0 + m01*in1 ; out1 = m10*in0 + m11*in1   (reference highperfkernels.cpp:336-343)
template <typename T>
__device__ __forceinline__ void apply2x2(const Mat2<T> &m, Cx<T> &a0, Cx<T> &a1) {
  Cx<T> in0 = a0, in1 = a1;
  a0 = cadd(cmul(m.m00, in0), cmul(m.m01, in1));
  a1 = cadd(cmul(m.m10, in0), cmul(m.m11, in1));
}

// The same update with contracted multiply-adds (IQSB_ARITH_FMA): 16 instead of 28 instructions.
__device__ __forceinline__ double fma_c(double a, double b, double c) { return fma(a, b, c); }
__device__ __forceinline__ float fma_c(float a, float b, float c) { return fmaf(a, b, c); }
template <typename T>
__device__ __forceinline__ void apply2x2_fma(const Mat2<T> &m, Cx<T> &a0, Cx<T> &a1) {
  const Cx<T> x = a0, y = a1;
  a0.re = fma_c(m.m00.re, x.re, fma_c(-m.m00.im, x.im, fma_c(m.m01.re, y.re, -m.m01.im * y.im)));
  a0.im = fma_c(m.m00.re, x.im, fma_c(m.m00.im, x.re, fma_c(m.m01.re, y.im, m.m01.im * y.re)));
  a1.re = fma_c(m.m10.re, x.re, fma_c(-m.m10.im, x.im, fma_c(m.m11.re, y.re, -m.m11.im * y.im)));
  a1.im = fma_c(m.m10.re, x.im, fma_c(m.m10.im, x.re, fma_c(m.m11.re, y.im, m.m11.im * y.re)));
}

// A chunk: two adjacent amplitudes.
template <typename T>
struct __align__(sizeof(T) * 4) Chunk {
  Cx<T> a, b;
};

// 256-bit (double) / 128-bit (float) global accesses.  `volatile` + "memory" keep the
// compiler from reordering them across each other (the kernels update in place).
__device__ __forceinline__ Chunk<double> ld_chunk(const Chunk<double> *p) {
  Chunk<double> r;
  asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];"
               : "=d"(r.a.re), "=d"(r.a.im), "=d"(r.b.re), "=d"(r.b.im)
               : "l"(p)
               : "memory");
  return r;
}
__device__ __forceinline__ void st_chunk(Chunk<double> *p, const Chunk<double> &r) {
  asm volatile("st.global.v4.f64 [%4], {%0,%1,%2,%3};" ::"d"(r.a.re), "d"(r.a.im), "d"(r.b.re),
               "d"(r.b.im), "l"(p)
               : "memory");
}
__device__ __forceinline__ Chunk<float> ld_chunk(const Chunk<float> *p) {
  Chunk<float> r;
  asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.a.re), "=f"(r.a.im), "=f"(r.b.re), "=f"(r.b.im)
               : "l"(p)
               : "memory");
  return r;
}
__device__ __forceinline__ void st_chunk(Chunk<float> *p, const Chunk<float> &r) {
  asm volatile("st.global.v4.f32 [%4], {%0,%1,%2,%3};" ::"f"(r.a.re), "f"(r.a.im), "f"(r.b.re),
               "f"(r.b.im), "l"(p)
               : "memory");
}
__device__ __forceinline__ Cx<double> ld_amp(const Cx<double> *p) {
  Cx<double> r;
  asm volatile("ld.global.v2.f64 {%0,%1}, [%2];" : "=d"(r.re), "=d"(r.im) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ void st_amp(Cx<double> *p, const Cx<double> &r) {
  asm volatile("st.global.v2.f64 [%2], {%0,%1};" ::"d"(r.re), "d"(r.im), "l"(p) : "memory");
}
__device__ __forceinline__ Cx<float> ld_amp(const Cx<float> *p) {
  Cx<float> r;
  asm volatile("ld.global.v2.f32 {%0,%1}, [%2];" : "=f"(r.re), "=f"(r.im) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ void st_amp(Cx<float> *p, const Cx<float> &r) {
  asm volatile("st.global.v2.f32 [%2], {%0,%1};" ::"f"(r.re), "f"(r.im), "l"(p) : "memory");
}

// Index geometry shared by every "subset of the index space" kernel: work item t in
// [0, nwork) is expanded by inserting a zero bit at each of ins[0] < ins[1] < ins[2]
// (63 = unused slot), in units of the kernel's access width (amplitudes or chunks).
struct Geom {
  uint64_t nwork;
  unsigned ins0, ins1, ins2;
  uint64_t off0, off1;  // added to the expanded index for the two partners
};

// (an unused Geom slot carries p = 63: the two-step shift keeps every shift count below 64)
__device__ __forceinline__ uint64_t insert_zero(uint64_t x, unsigned p) {
  uint64_t low = x & ((1ull << p) - 1ull);
  return (((x >> p) << p) << 1) | low;
}
__device__ __forceinline__ uint64_t expand(uint64_t t, const Geom &g) {
  uint64_t x = insert_zero(t, g.ins0);
  x = insert_zero(x, g.ins1);
  x = insert_zero(x, g.ins2);
  return x;
}

#endif  // __CUDACC__

// ---------------------------------------------------------------------------------------
// launch helpers (host)
// ---------------------------------------------------------------------------------------
static inline Geom make_geom(uint64_t nwork, int nins, const unsigned *ins_sorted, uint64_t off0,
                             uint64_t off1) {
  Geom g;
  g.nwork = nwork;
  g.ins0 = nins > 0 ? ins_sorted[0] : 63u;
  g.ins1 = nins > 1 ? ins_sorted[1] : 63u;
  g.ins2 = nins > 2 ? ins_sorted[2] : 63u;
  g.off0 = off0;
  g.off1 = off1;
  return g;
}

template <typename T>
static inline Mat2<T> make_mat(const double m[8]) {
  Mat2<T> r;
  r.m00 = {(T)m[0], (T)m[1]};
  r.m01 = {(T)m[2], (T)m[3]};
  r.m10 = {(T)m[4], (T)m[5]};
  r.m11 = {(T)m[6], (T)m[7]};
  return r;
}

// after every kernel launch: error check, launch count and (when profiling) a CUDA event on the stream.
// `what` must be a string literal; `algo_bytes` = algorithmic bytes of the launch (0: not stated)
int iqsb_check_launch(iqsb_ctx *ctx, const char *what, double algo_bytes = 0.0);

// kernels_gate.cu
int iqsb_launch_pairs(iqsb_state *st, void *s0, void *s1, int width, const Geom &g, const double m[8]);
int iqsb_launch_inchunk(iqsb_state *st, void *s, const Geom &g, const double m[8]);
int iqsb_launch_scale_subset(iqsb_state *st, void *s, int width, const Geom &g, const double f[2]);
int iqsb_launch_zero_subset(iqsb_state *st, void *s, int width, const Geom &g);
