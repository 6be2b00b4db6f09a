// kernels_fused.cu -- gate fusion: a batch of gates applied in as few sweeps over HBM as
// possible, the amplitudes staged in shared-memory tiles and worked on in registers.
//
// Replaces ApplyFusedGates (reference src/qureg_fusion.cpp:55-94), which replays the queued gates
// block by block, with blocks of 2^log2llc CONTIGUOUS amplitudes sized for the CPU's last-level
// cache -- so only gates whose target lies below log2llc can be fused there.
//
// Three levels:
//
//   run    a tile is the set of 2^K amplitudes (K = 12: 64 KiB of ComplexDP) whose indices differ
//          only in K chosen bit positions: the four lowest positions are always part of it (global
//          accesses are 256-byte runs), the other eight are whatever positions the gates of the run
//          act on.  A run costs ONE read and ONE write of the state, wherever its target qubits
//          sit.  Controls may be anywhere: outside the tile they select whole tiles (the reference's
//          rule for controls above the block, src/qureg_applyctrl1qubitgate.cpp:296-309).
//   group  inside a run, consecutive gates whose targets fall on at most THREE tile bits form a
//          group: every thread takes the 8 amplitudes spanned by those three bits into registers,
//          applies all gates of the group there and writes them back -- one shared-memory round
//          trip and one __syncthreads per group instead of per gate (round 1: per gate, 73 % of the
//          shared-memory wavefront budget).
//   gate   each gate is classified on the host by the zero structure of its matrix (general / real
//          / diagonal / diag(1,d) / anti-diagonal / exact X / real-diagonal+imaginary-off-diagonal
//          = RX) and the kernel switches once per gate, CTA-uniformly, to code that leaves out the
//          products with exact zeros and ones: 0-12 instead of 28 FP64 instructions per pair.  For
//          finite amplitudes the values are identical to the reference's full evaluation (a product
//          with an exact zero only contributes a signed zero); the reference does the same on the
//          CPU for named gates (src/spec_kernels.cpp:36-153).  sqrt X / sqrt Y (entries +-1/2 +- i/2) take
//          12 instructions: halving is exact and commutes with the rounding of the sums.
//   perm   an exact X / CNOT that shares no qubit with the gates after it in its group does not touch the
//          registers at all: a permutation of the 8 register-resident amplitudes is an affine map of the
//          3-bit register index over GF(2), composed on the host into the basis (w, c0) the write-back
//          addresses are built from; an X whose control is a thread bit or a bit of the tile's base index
//          is one conditional XOR on the address.  (As register moves these gates cost 5 ms per group and
//          2^32 amplitudes; a group of them now runs at the shared-memory bandwidth, 4.3 ms.)
//
// Descriptors (group headers, gates with their matrices) travel as a __grid_constant__ kernel parameter:
// they are read with warp-uniform indices through the constant cache and cost no shared-memory
// wavefronts -- the resource the tile phase is short of; only the per-lane slot tables are staged in
// shared memory.  One launch (one sweep) carries up to 24 groups / 48 gates.
//
// Planner (host, exposed as iqsb_plan_fused_order): runs are cut greedily in program order, but a
// pure-permutation gate (X / CNOT with an exact 0/1 matrix) commutes EXACTLY -- no rounding is
// involved -- with every gate on other qubits, so it is allowed to move across skipped gates into
// the earliest run that holds its target.  A layer of 32 one-qubit gates + 16 CNOTs is 4 sweeps
// instead of 6, bit for bit the same result.  Arithmetic gates never change their relative order.
//
// Shared-memory layout: 16-byte slots, slot index i stored at i ^ fold(i) where fold XORs the 3-bit
// fields i[5:3], i[8:6], i[11:9] into the low three bits.  The 8 lanes of a quarter-warp vary three
// tile bits chosen by the planner with distinct residues mod 3, which makes every access of a group
// (and the global<->tile copies) bank-conflict free whatever the group's register bits are.
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "iqsb_internal.cuh"

namespace {

// Two CTA shapes, picked per batch of gates (iqsb_fused): tiles of 2^12 amplitudes worked on by 256 threads,
// 3 CTAs per SM -- or tiles of 2^11 with 128 threads, 6 CTAs per SM, which interleave the phases of a run
// more finely (5-6 % faster per run, sweeps at 105 % of the copy peak) but hold one target position less.
// The smaller tile is used whenever it does not cost an extra run.
#ifndef IQSB_FUSED_TILE
#define IQSB_FUSED_TILE 12
#endif
#ifndef IQSB_FUSED_REGBITS
#define IQSB_FUSED_REGBITS 3
#endif
constexpr int kSmallTile = 11;
constexpr int kMaxFusedGates = 4096;
constexpr int kTile = IQSB_FUSED_TILE;  // tile exponent (<= 12)
// (A variant with two tile buffers of 2^11 amplitudes per CTA, the next tile fetched while the current
// one is worked on, was measured and dropped -- profiles/r02f_fused_variants.log: 68.7 vs 61.2 ms per
// bench layer on the same box.  With 3 CTAs per SM the loads of one CTA already overlap the
// arithmetic of the others; the kernel is bound by instruction issue inside the tile.)
constexpr int kLow = 4;                 // lowest positions always in the tile
constexpr int kRegBits = IQSB_FUSED_REGBITS;  // tile bits held in registers by a group (3 or 4)
constexpr int kAmps = 1 << kRegBits;          // amplitudes per thread
static_assert(kRegBits == 3 || kRegBits == 4, "register blocking");
constexpr int kBatchGates = 48;         // descriptors resident in shared memory at a time
constexpr int kBatchGroups = 24;
constexpr int kReorderWindow = 512;     // how far the planner looks past the first skipped gate
static_assert(kTile >= 9 && kTile <= 12, "tile geometry");

enum GateClass : uint8_t { kGeneral = 0, kReal, kDiag, kDiag1, kAnti, kXExact, kRealDiagImagOff, kSqrtX, kSqrtY };

template <typename T>
struct alignas(16) FGate {
  Mat2<T> m;
  uint8_t cls;    // GateClass
  uint8_t tbit;   // register bit of the target (0..2)
  uint8_t ckind;  // 0 none | 1 register bit (pairs enabled: `en`) | 2 thread bit `c` | 3 bit `c` of the tile's base index
  uint8_t c;
  uint8_t en;     // mask of the kAmps/2 register pairs the gate acts on
  uint8_t last;   // 1: last gate of the group's main section
  uint8_t trail;  // 0: applied to the registers | 1: exact X whose control is a thread / tile bit, folded into the
                  //    write-back address (`pu`) | 2: exact X / CNOT absorbed into the group's write-back basis
  uint8_t pad8;
  uint32_t origin;  // index of the gate in the caller's list (iqsb_plan_fused_trace)
  uint16_t pu;      // trail == 1: swizzled slot offset XORed into the write-back address when the control is set
  uint16_t pad16;
};

struct alignas(16) GroupDesc {
  uint16_t lo[32];  // swizzled slot contributed by thread bits 0..4
  uint16_t hi[16];  // ... by thread bits 5..8
  uint16_t p[4];    // swizzled slot offset of register bit k (where the registers are loaded from)
  uint16_t gate_first, gate_count;  // all gates of the group: main section, then trail == 1, then trail == 2
  uint16_t log2_threads;  // tile exponent - kRegBits
  uint16_t nmain, ncond;  // gates applied to the registers / conditional write-back offsets
  uint16_t w[4];          // write-back basis: register r goes to slot px ^ c0 ^ XOR_j r_j w[j]
  uint16_t c0;
  uint16_t pad[2];
};
static_assert(sizeof(GroupDesc) == 128, "group descriptor layout");

struct BatchHdr {
  int ngroups, ngates, pad0, pad1;
};

// Descriptors of one launch in the kernel's parameter space (constant bank): group headers and gates are
// read with warp-uniform indices through the constant cache -- they cost no shared-memory wavefronts, which
// is what the tile phase is short of -- and only the per-lane slot tables go to shared memory.
struct GroupHdr {
  uint16_t p[4];
  uint16_t w[4];
  uint16_t gate_first, nmain, ncond, log2_threads, c0, pad[3];
};
static_assert(sizeof(GroupHdr) == 32, "group header layout");
template <typename T>
struct RunParams {
  int ngroups, pad0, pad1, pad2;
  GroupHdr hdr[kBatchGroups];
  uint16_t lohi[kBatchGroups][48];  // lo[32] then hi[16] of every group
  FGate<T> gates[kBatchGates];
};

template <typename T>
__host__ __device__ constexpr size_t batch_stride() {
  return sizeof(BatchHdr) + kBatchGroups * sizeof(GroupDesc) + kBatchGates * sizeof(FGate<T>);
}

struct TileDesc {
  uint8_t pos[kTile];
  int nS;
};

// 16-byte slots, XOR-folded swizzle (linear over GF(2): swz(a ^ b) == swz(a) ^ swz(b))
__host__ __device__ __forceinline__ unsigned swz(unsigned i) { return i ^ (((i >> 3) ^ (i >> 6) ^ (i >> 9)) & 7u); }

// global -> tile without register staging: one asynchronous 16-byte (ComplexDP) / 8-byte (ComplexSP)
// copy per amplitude, straight into its swizzled slot (LDGSTS); every thread has its whole share of
// the tile in flight at once.
__device__ __forceinline__ void cp_async_amp(Cx<double> *smem, const Cx<double> *gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_amp(Cx<float> *smem, const Cx<float> *gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gmem) : "memory");
}
template <typename T, int kThreads>
__device__ __forceinline__ void tile_load_async(Cx<T> *tile, const Chunk<T> *g, const uint32_t *g_lo, const uint32_t *g_hi, unsigned nchunks) {
#pragma unroll 4
  for (unsigned c = threadIdx.x; c < nchunks; c += kThreads) {
    const Cx<T> *src = reinterpret_cast<const Cx<T> *>(g + ((uint64_t)g_lo[c & 255] | (uint64_t)g_hi[c >> 8]));
    const unsigned s = swz(2 * c);
    cp_async_amp(tile + s, src);
    cp_async_amp(tile + (s ^ 1u), src + 1);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
// tile -> global, U 32-byte stores in flight per thread; nchunks is a multiple of kThreads * U
template <typename T, int U, int kThreads>
__device__ __forceinline__ void tile_store(const Cx<T> *tile, Chunk<T> *g, const uint32_t *g_lo, const uint32_t *g_hi, unsigned nchunks) {
#pragma unroll 1
  for (unsigned c0 = threadIdx.x; c0 < nchunks; c0 += kThreads * U) {
    Chunk<T> v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned s = swz(2 * (c0 + u * kThreads));
      v[u].a = tile[s];
      v[u].b = tile[s ^ 1u];
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      unsigned c = c0 + u * kThreads;
      st_chunk(g + ((uint64_t)g_lo[c & 255] | (uint64_t)g_hi[c >> 8]), v[u]);
    }
  }
}

// ---- one gate on the 8 register-resident amplitudes -----------------------------------------
// B = register bit of the target; pair k (k = 0..3) is (a[r0], a[r0 | 1 << B]) with r0 = k with a
// zero inserted at bit B.  `en` selects the pairs (control on another register bit).
template <typename T, int B, typename F>
__device__ __forceinline__ void for_pairs(Cx<T> (&a)[kAmps], F f) {
#pragma unroll
  for (int k = 0; k < kAmps / 2; ++k) {
    const int r0 = ((k >> B) << (B + 1)) | (k & ((1 << B) - 1));
    f(a[r0], a[r0 | (1 << B)]);
  }
}

template <typename T, bool FMA, int B>
__device__ __forceinline__ void apply_on_bit(unsigned cls, unsigned en, const Mat2<T> &m, Cx<T> (&a)[kAmps]) {
  switch (cls) {
    case kXExact:  // out0 = in1, out1 = in0.  The only class that takes a control among the register bits
                   // (`en` = the pairs whose control bit is set): CNOTs are moves between registers.
      if (en == (1u << (kAmps / 2)) - 1u) {  // plain X: every pair trades places
        for_pairs<T, B>(a, [&](Cx<T> &x, Cx<T> &y) {
          const Cx<T> t = x;
          x = y;
          y = t;
        });
        break;
      }
#pragma unroll
      for (int k = 0; k < kAmps / 2; ++k) {
        const int r0 = ((k >> B) << (B + 1)) | (k & ((1 << B) - 1));
        const bool on = (en >> k) & 1u;
        const Cx<T> x = a[r0], y = a[r0 | (1 << B)];
        a[r0].re = on ? y.re : x.re;
        a[r0].im = on ? y.im : x.im;
        a[r0 | (1 << B)].re = on ? x.re : y.re;
        a[r0 | (1 << B)].im = on ? x.im : y.im;
      }
      break;
    case kDiag1: {  // out0 = in0, out1 = m11 * in1
      const Cx<T> d = m.m11;
      for_pairs<T, B>(a, [&](Cx<T> &, Cx<T> &y) { y = cmul(d, y); });
      break;
    }
    case kDiag: {
      const Cx<T> d0 = m.m00, d1 = m.m11;
      for_pairs<T, B>(a, [&](Cx<T> &x, Cx<T> &y) {
        x = cmul(d0, x);
        y = cmul(d1, y);
      });
      break;
    }
    case kAnti: {
      const Cx<T> u = m.m01, l = m.m10;
      for_pairs<T, B>(a, [&](Cx<T> &x, Cx<T> &y) {
        const Cx<T> t = cmul(u, y);
        y = cmul(l, x);
        x = t;
      });
      break;
    }
    case kReal: {  // every imaginary part of the matrix is an exact zero (H, RY)
      const T r00 = m.m00.re, r01 = m.m01.re, r10 = m.m10.re, r11 = m.m11.re;
      for_pairs<T, B>(a, [&](Cx<T> &x, Cx<T> &y) {
        Cx<T> o0, o1;
        if (FMA) {
          o0.re = fma_c(r00, x.re, r01 * y.re);
          o0.im = fma_c(r00, x.im, r01 * y.im);
          o1.re = fma_c(r10, x.re, r11 * y.re);
          o1.im = fma_c(r10, x.im, r11 * y.im);
        } else {
          o0.re = add_rn(mul_rn(r00, x.re), mul_rn(r01, y.re));
          o0.im = add_rn(mul_rn(r00, x.im), mul_rn(r01, y.im));
          o1.re = add_rn(mul_rn(r10, x.re), mul_rn(r11, y.re));
          o1.im = add_rn(mul_rn(r10, x.im), mul_rn(r11, y.im));
        }
        x = o0;
        y = o1;
      });
      break;
    }
    case kRealDiagImagOff: {  // m00, m11 real; m01, m10 imaginary (RX)
      const T r00 = m.m00.re, i01 = m.m01.im, i10 = m.m10.im, r11 = m.m11.re;
      for_pairs<T, B>(a, [&](Cx<T> &x, Cx<T> &y) {
        Cx<T> o0, o1;
        if (FMA) {
          o0.re = fma_c(r00, x.re, -(i01 * y.im));
          o0.im = fma_c(r00, x.im, i01 * y.re);
          o1.re = fma_c(r11, y.re, -(i10 * x.im));
          o1.im = fma_c(r11, y.im, i10 * x.re);
        } else {
          o0.re = sub_rn(mul_rn(r00, x.re), mul_rn(i01, y.im));
          o0.im = add_rn(mul_rn(r00, x.im), mul_rn(i01, y.re));
          o1.re = add_rn(-mul_rn(i10, x.im), mul_rn(r11, y.re));
          o1.im = add_rn(mul_rn(i10, x.re), mul_rn(r11, y.im));
        }
        x = o0;
        y = o1;
      });
      break;
    }
    case kSqrtX:
    case kSqrtY: {
      // every entry is +-1/2 +- i/2 (reference src/qureg_apply1qubitgate.cpp:389-431).  A product with 1/2 is
      // exact and fl(a/2 - b/2) = fl(a - b)/2, so the reference's 28 operations collapse to 4 sums of the
      // inputs, 4 sums of those and 4 halvings -- the same values unless an intermediate is subnormal.
      const bool sx = cls == kSqrtX;
      for_pairs<T, B>(a, [&](Cx<T> &x, Cx<T> &y) {
        const T A = sub_rn(x.re, x.im), Bp = add_rn(x.re, x.im), Cp = add_rn(y.re, y.im), D = sub_rn(y.re, y.im);
        const T h = (T)0.5;
        Cx<T> o0, o1;
        if (sx) {  // 1/2 [[1+i, 1-i], [1-i, 1+i]]
          o0.re = mul_rn(h, add_rn(A, Cp));
          o0.im = mul_rn(h, sub_rn(Bp, D));
          o1.re = mul_rn(h, add_rn(Bp, D));
          o1.im = mul_rn(h, sub_rn(Cp, A));
        } else {   // 1/2 [[1+i, -1-i], [1+i, 1+i]]
          o0.re = mul_rn(h, sub_rn(A, D));
          o0.im = mul_rn(h, sub_rn(Bp, Cp));
          o1.re = mul_rn(h, add_rn(A, D));
          o1.im = mul_rn(h, add_rn(Bp, Cp));
        }
        x = o0;
        y = o1;
      });
      break;
    }
    default:
      for_pairs<T, B>(a, [&](Cx<T> &x, Cx<T> &y) {
        if (FMA) apply2x2_fma(m, x, y);
        else apply2x2(m, x, y);
      });
      break;
  }
}

template <typename T, bool FMA, int kThreads>
__global__ void __launch_bounds__(kThreads, 768 / kThreads)
    k_fused(Chunk<T> *__restrict__ state, uint64_t nouter, TileDesc td, unsigned long long *__restrict__ next_tile, int debug_no_io,
              const __grid_constant__ RunParams<T> P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Cx<T> *tile = reinterpret_cast<Cx<T> *>(smem_raw);
  __shared__ uint32_t g_lo[256], g_hi[8];
  __shared__ uint16_t s_lohi[kBatchGroups][48];
  __shared__ int s_pos[16];
  __shared__ unsigned long long s_tile;
  const int nS = td.nS;  // pos[0] == 0 always
  if (threadIdx.x < kTile) s_pos[threadIdx.x] = td.pos[threadIdx.x];
  for (int i = threadIdx.x; i < P.ngroups * 48; i += kThreads) s_lohi[i / 48][i % 48] = P.lohi[i / 48][i % 48];
  __syncthreads();
  for (unsigned t = threadIdx.x; t < 256 + 8; t += kThreads) {
    unsigned v = t < 256 ? t : (t - 256) << 8;
    uint32_t go = 0;
#pragma unroll 1
    for (int k = 1; k < nS; ++k)
      if ((v >> (k - 1)) & 1u) go |= 1u << (s_pos[k] - 1);
    if (t < 256) g_lo[t] = go;
    else g_hi[t - 256] = go;
  }
  __syncthreads();
  const unsigned nchunks = 1u << (nS - 1);
  const int ngroups = P.ngroups;
  for (unsigned it = 0;; ++it) {
    if (threadIdx.x == 0) {
      uint64_t o = next_tile != nullptr ? atomicAdd(next_tile, 1ull) : (uint64_t)blockIdx.x + (uint64_t)it * gridDim.x;
      uint64_t base = ~0ull;
      if (o < nouter) {
        base = o;
#pragma unroll 1
        for (int k = 0; k < nS; ++k) base = insert_zero(base, (unsigned)s_pos[k]);
      }
      s_tile = base;
    }
    __syncthreads();
    if (s_tile == ~0ull) break;
    if (!debug_no_io) {
      tile_load_async<T, kThreads>(tile, state + (s_tile >> 1), g_lo, g_hi, nchunks);
      cp_async_wait<0>();
    }
    __syncthreads();  // the tile is loaded
#pragma unroll 1
    for (int gi = 0; gi < ngroups; ++gi) {
      const GroupHdr &H = P.hdr[gi];
      auto slots = [&](unsigned px, const uint16_t (&basis)[4], unsigned (&sl)[kAmps]) {
        unsigned Pq[kRegBits];
#pragma unroll
        for (int j = 0; j < kRegBits; ++j) Pq[j] = basis[j];
#pragma unroll
        for (int r = 0; r < kAmps; ++r) {
          unsigned sidx = px;
#pragma unroll
          for (int j = 0; j < kRegBits; ++j)
            if (r & (1 << j)) sidx ^= Pq[j];
          sl[r] = sidx;
        }
      };
      const unsigned nthr = 1u << H.log2_threads;
#pragma unroll 1
      for (unsigned t = threadIdx.x; t < nthr; t += kThreads) {
        Cx<T> a[kAmps];
        const unsigned px = (unsigned)s_lohi[gi][t & 31u] ^ (unsigned)s_lohi[gi][32 + (t >> 5)];
        {
          unsigned sl[kAmps];
          slots(px, H.p, sl);
#pragma unroll
          for (int r = 0; r < kAmps; ++r) a[r] = tile[sl[r]];
        }
        bool more = H.nmain != 0;
#pragma unroll 1
        for (int gj = H.gate_first; more; ++gj) {
          const FGate<T> &fg = P.gates[gj];
          const unsigned cls = fg.cls, tbit = fg.tbit, ckind = fg.ckind, c = fg.c, en = fg.en;
          more = fg.last == 0;
          if (ckind == 3 && !((s_tile >> c) & 1ull)) continue;  // uniform over the CTA
          if (ckind == 2 && !((t >> c) & 1u)) continue;       // uniform over the warp when c >= 5 (the planner's choice)
          if (tbit == 0) apply_on_bit<T, FMA, 0>(cls, en, fg.m, a);
          else if (tbit == 1) apply_on_bit<T, FMA, 1>(cls, en, fg.m, a);
          else if (kRegBits == 3 || tbit == 2) apply_on_bit<T, FMA, 2>(cls, en, fg.m, a);
          else apply_on_bit<T, FMA, kRegBits - 1>(cls, en, fg.m, a);
        }
        {
          // Exact X / CNOT gates at the end of a group never touch the registers: a permutation of the
          // 8 amplitudes is a change of the basis (w, c0) the write-back addresses are built from, and an
          // X whose control is a thread bit or a bit of the tile's base index is one more conditional offset.
          unsigned pxs = px ^ H.c0;
#pragma unroll 1
          for (int gj = H.gate_first + H.nmain, ge = gj + H.ncond; gj < ge; ++gj) {
            const FGate<T> &fg = P.gates[gj];
            const unsigned c = fg.c;
            const bool on = fg.ckind == 2 ? ((t >> c) & 1u) != 0u : ((s_tile >> c) & 1ull) != 0ull;
            if (on) pxs ^= fg.pu;
          }
          unsigned sl[kAmps];
          slots(pxs, H.w, sl);
#pragma unroll
          for (int r = 0; r < kAmps; ++r) tile[sl[r]] = a[r];
        }
      }
      __syncthreads();
    }
    Chunk<T> *g = state + (s_tile >> 1);
    if (debug_no_io) {
    } else if (nchunks % (kThreads * 2) == 0) tile_store<T, 2, kThreads>(tile, g, g_lo, g_hi, nchunks);
    else tile_store<T, 1, kThreads>(tile, g, g_lo, g_hi, nchunks);
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// host: classification, groups, descriptors
// ---------------------------------------------------------------------------------------------
bool is_xexact(const double m[8]) {
  return m[0] == 0. && m[1] == 0. && m[6] == 0. && m[7] == 0. && m[2] == 1. && m[3] == 0. && m[4] == 1. && m[5] == 0.;
}

uint8_t classify(const double m[8]) {
  const bool z00 = m[0] == 0. && m[1] == 0., z01 = m[2] == 0. && m[3] == 0., z10 = m[4] == 0. && m[5] == 0., z11 = m[6] == 0. && m[7] == 0.;
  if (z00 && z11) return is_xexact(m) ? kXExact : kAnti;
  if (z01 && z10) return (m[0] == 1. && m[1] == 0.) ? kDiag1 : kDiag;
  if (m[1] == 0. && m[3] == 0. && m[5] == 0. && m[7] == 0.) return kReal;
  if (m[1] == 0. && m[7] == 0. && m[2] == 0. && m[4] == 0.) return kRealDiagImagOff;
  static const double sqrt_x[8] = {0.5, 0.5, 0.5, -0.5, 0.5, -0.5, 0.5, 0.5}, sqrt_y[8] = {0.5, 0.5, -0.5, -0.5, 0.5, 0.5, 0.5, 0.5};
  if (memcmp(m, sqrt_x, sizeof(sqrt_x)) == 0) return kSqrtX;
  if (memcmp(m, sqrt_y, sizeof(sqrt_y)) == 0) return kSqrtY;
  return kGeneral;
}

struct HostGroup {
  std::vector<int> gates;  // indices into the run's gate list
  int rs[kRegBits];        // register slots (tile-local bits)
  int nrs = 0;
  uint64_t qmask = 0;  // positions (targets and controls) its gates touch
  unsigned arith_ctrl = 0;  // tile slots that control an ARITHMETIC gate of the group: never register bits
  bool has(int slot) const {
    for (int j = 0; j < nrs; ++j)
      if (rs[j] == slot) return true;
    return false;
  }
};

// Build the batches (header + groups + gates) of one run.  `run` lists indices into `in`, in
// execution order; every target is in the tile.
template <typename T>
void build_batches(const iqsb_fgate *in, const std::vector<int> &run, const TileDesc &td, bool reorder, std::vector<unsigned char> &out, int &nbatches) {
  const int nS = td.nS;
  int slot_of[64];
  for (int p = 0; p < 64; ++p) slot_of[p] = -1;
  for (int k = 0; k < nS; ++k) slot_of[td.pos[k]] = k;
  // 1. groups: gates on at most kRegBits distinct target slots.  Arithmetic gates join the last
  //    group (or open a new one); an exact X / CNOT commutes without rounding with gates on other
  //    qubits, so it may join the EARLIEST group after the last gate it shares a qubit with.
  std::vector<HostGroup> groups;
  for (size_t k = 0; k < run.size(); ++k) {
    const iqsb_fgate &q = in[run[k]];
    const int ts = slot_of[q.target];
    const uint64_t qm = (1ull << (unsigned)q.target) | (q.kind == 1 ? 1ull << (unsigned)q.control : 0ull);
    // Only exact X takes a control among the register bits (there it is a move between registers);
    // the control of an arithmetic gate must be a thread bit, so that the arithmetic classes stay
    // straight-line code over all register pairs.
    const bool perm = is_xexact(q.m);
    const int cs = q.kind == 1 ? slot_of[q.control] : -1;
    const bool arith_ctrl = !perm && cs >= 0;
    auto room = [&](const HostGroup &g) {
      if (!(g.has(ts) || g.nrs < kRegBits) || (int)g.gates.size() >= kBatchGates) return false;
      if (!g.has(ts) && ((g.arith_ctrl >> ts) & 1u)) return false;  // ts would become a register bit
      if (arith_ctrl) {
        if (g.has(cs)) return false;
        // enough other slots must remain to fill the register bits
        const unsigned ac = g.arith_ctrl | (1u << cs);
        if (__builtin_popcount(ac) + kRegBits > nS) return false;
      }
      return true;
    };
    int where = -1;
    if (reorder && perm && !groups.empty()) {
      int e = 0;
      for (int j = (int)groups.size() - 1; j >= 0; --j)
        if (groups[j].qmask & qm) { e = j; break; }
      for (int j = e; j < (int)groups.size() && where < 0; ++j)
        if (room(groups[j])) where = j;
    } else if (!groups.empty() && room(groups.back())) {
      where = (int)groups.size() - 1;
    }
    if (where < 0) {
      groups.push_back(HostGroup());
      where = (int)groups.size() - 1;
    }
    HostGroup &g = groups[where];
    if (!g.has(ts)) g.rs[g.nrs++] = ts;
    g.gates.push_back((int)k);
    g.qmask |= qm;
    if (arith_ctrl) g.arith_ctrl |= 1u << cs;
  }
  // 2. descriptors
  out.clear();
  nbatches = 0;
  BatchHdr hdr = {0, 0, 0, 0};
  std::vector<GroupDesc> bgroups;
  std::vector<FGate<T>> bgates;
  auto flush = [&]() {
    if (bgroups.empty()) return;
    hdr.ngroups = (int)bgroups.size();
    hdr.ngates = (int)bgates.size();
    size_t off = out.size();
    out.resize(off + batch_stride<T>(), 0);
    memcpy(&out[off], &hdr, sizeof(hdr));
    memcpy(&out[off + sizeof(BatchHdr)], bgroups.data(), bgroups.size() * sizeof(GroupDesc));
    memcpy(&out[off + sizeof(BatchHdr) + kBatchGroups * sizeof(GroupDesc)], bgates.data(), bgates.size() * sizeof(FGate<T>));
    bgroups.clear();
    bgates.clear();
    ++nbatches;
  };
  for (HostGroup &hg : groups) {
    if ((int)bgroups.size() == kBatchGroups || bgates.size() + hg.gates.size() > (size_t)kBatchGates) flush();
    bool used[16] = {false};
    for (int j = 0; j < hg.nrs; ++j) used[hg.rs[j]] = true;
    // tile-local controls of the group
    bool is_ctrl[16] = {false};
    for (int k : hg.gates) {
      const iqsb_fgate &q = in[run[k]];
      if (q.kind == 1 && slot_of[q.control] >= 0) is_ctrl[slot_of[q.control]] = true;
    }
    // spare register bits: controls of X gates first (a CNOT is then a move between registers), then
    // any slot that does not control an arithmetic gate
    for (int s = 0; s < nS && hg.nrs < kRegBits; ++s)
      if (is_ctrl[s] && !used[s] && !((hg.arith_ctrl >> s) & 1u)) { hg.rs[hg.nrs++] = s; used[s] = true; }
    for (int s = nS - 1; s >= 0 && hg.nrs < kRegBits; --s)
      if (!used[s] && !((hg.arith_ctrl >> s) & 1u)) { hg.rs[hg.nrs++] = s; used[s] = true; }
    // thread bits -> tile slots.  Bits 0..2 (the lanes of a quarter-warp): three slots with distinct
    // residues mod 3, preferably not controls; controls go to the highest thread bits (warp-uniform).
    int dep[16], ndep = 0;
    bool taken[16] = {false};
    for (int pass = 0; pass < 2; ++pass)
      for (int res = 0; res < 3; ++res) {
        bool have = false;
        for (int j = 0; j < ndep; ++j) have = have || dep[j] % 3 == res;
        if (have || ndep >= 3) continue;
        for (int s = 0; s < nS; ++s)
          if (!used[s] && !taken[s] && s % 3 == res && (pass == 1 || !is_ctrl[s])) { dep[ndep++] = s; taken[s] = true; break; }
      }
    for (int s = 0; s < nS; ++s)
      if (!used[s] && !taken[s] && !is_ctrl[s]) { dep[ndep++] = s; taken[s] = true; }
    for (int s = 0; s < nS; ++s)
      if (!used[s] && !taken[s]) { dep[ndep++] = s; taken[s] = true; }
    int tbit_of[16];
    for (int s = 0; s < 16; ++s) tbit_of[s] = -1;
    for (int j = 0; j < ndep; ++j) tbit_of[dep[j]] = j;
    GroupDesc gd;
    memset(&gd, 0, sizeof(gd));
    for (unsigned v = 0; v < 32; ++v) {
      unsigned x = 0;
      for (int k = 0; k < 5 && k < ndep; ++k)
        if ((v >> k) & 1u) x |= 1u << dep[k];
      gd.lo[v] = (uint16_t)swz(x);
    }
    for (unsigned v = 0; v < 16; ++v) {
      unsigned x = 0;
      for (int k = 5; k < ndep; ++k)
        if ((v >> (k - 5)) & 1u) x |= 1u << dep[k];
      gd.hi[v] = (uint16_t)swz(x);
    }
    for (int j = 0; j < kRegBits; ++j) gd.p[j] = (uint16_t)swz(1u << hg.rs[j]);
    auto regbit = [&](int slot) {
      for (int j = 0; j < kRegBits; ++j)
        if (hg.rs[j] == slot) return j;
      return -1;
    };
    // Which exact X / CNOT gates can be taken out of the arithmetic: walking backwards, a permutation gate
    // that shares no qubit with any gate staying behind it commutes with those exactly, so it can run last
    // -- and a permutation of the 8 register-resident amplitudes run last is only a different write-back
    // address.  trail[k]: 0 main, 1 conditional offset (control on a thread / tile bit), 2 absorbed.
    std::vector<uint8_t> trail(hg.gates.size(), 0);
    {
      uint64_t main_q = 0;
      for (int k = (int)hg.gates.size() - 1; k >= 0; --k) {
        const iqsb_fgate &q = in[run[hg.gates[k]]];
        const uint64_t qm = (1ull << (unsigned)q.target) | (q.kind == 1 ? 1ull << (unsigned)q.control : 0ull);
        if (reorder && is_xexact(q.m) && !(qm & main_q)) {
          const int cs = q.kind == 1 ? slot_of[q.control] : -1;
          trail[k] = (q.kind == 1 && !(cs >= 0 && regbit(cs) >= 0)) ? 1 : 2;
        } else {
          main_q |= qm;
        }
      }
    }
    // the trailing gates as one affine map of the 3-bit register index: T(q) = A q + v0 + sum_i ctrl_i u_i
    unsigned A[kRegBits], v0 = 0;
    for (int j = 0; j < kRegBits; ++j) A[j] = 1u << j;
    std::vector<std::pair<int, unsigned>> cond;  // (index into hg.gates, u)
    for (size_t k = 0; k < hg.gates.size(); ++k) {
      if (!trail[k]) continue;
      const iqsb_fgate &q = in[run[hg.gates[k]]];
      const int bt = regbit(slot_of[q.target]);
      if (trail[k] == 1) {
        cond.push_back({(int)k, 1u << bt});
      } else if (q.kind == 0) {
        v0 ^= 1u << bt;
      } else {  // CNOT between two register bits: M = I + e_b e_c^T applied to everything so far
        const int bc = regbit(slot_of[q.control]);
        for (int j = 0; j < kRegBits; ++j)
          if ((A[j] >> bc) & 1u) A[j] ^= 1u << bt;
        if ((v0 >> bc) & 1u) v0 ^= 1u << bt;
        for (auto &cu : cond)
          if ((cu.second >> bc) & 1u) cu.second ^= 1u << bt;
      }
    }
    auto span = [&](unsigned bits) {  // XOR of the load offsets of the register bits in `bits`
      unsigned o = 0;
      for (int j = 0; j < kRegBits; ++j)
        if ((bits >> j) & 1u) o ^= gd.p[j];
      return o;
    };
    for (int j = 0; j < kRegBits; ++j) gd.w[j] = (uint16_t)span(A[j]);
    gd.c0 = (uint16_t)span(v0);
    gd.gate_first = (uint16_t)bgates.size();
    gd.gate_count = (uint16_t)hg.gates.size();
    gd.log2_threads = (uint16_t)ndep;
    auto emit = [&](size_t k, uint8_t tr, unsigned pu) {
      const iqsb_fgate &q = in[run[hg.gates[k]]];
      FGate<T> o;
      memset(&o, 0, sizeof(o));
      o.m = make_mat<T>(q.m);
      o.cls = classify(q.m);
      o.origin = (uint32_t)run[hg.gates[k]];
      o.trail = tr;
      o.pu = (uint16_t)pu;
      const int ts = slot_of[q.target];
      o.tbit = (uint8_t)regbit(ts);
      o.en = (uint8_t)((1u << (kAmps / 2)) - 1u);
      if (q.kind == 1) {
        const int cs = slot_of[q.control];
        if (cs < 0) { o.ckind = 3; o.c = (uint8_t)q.control; }
        else if (tbit_of[cs] >= 0) { o.ckind = 2; o.c = (uint8_t)tbit_of[cs]; }
        else {
          const int cb = regbit(cs);
          o.ckind = 1;
          o.c = (uint8_t)cb;
          o.en = 0;
          for (int kk = 0; kk < kAmps / 2; ++kk) {
            const int r0 = ((kk >> o.tbit) << (o.tbit + 1)) | (kk & ((1 << o.tbit) - 1));
            if ((r0 >> cb) & 1) o.en |= (uint8_t)(1u << kk);
          }
        }
      }
      bgates.push_back(o);
    };
    int nmain = 0;
    for (size_t k = 0; k < hg.gates.size(); ++k)
      if (!trail[k]) { emit(k, 0, 0); ++nmain; }
    if (nmain) bgates.back().last = 1;
    for (auto &cu : cond) emit((size_t)cu.first, 1, span(cu.second));
    for (size_t k = 0; k < hg.gates.size(); ++k)
      if (trail[k] == 2) emit(k, 2, 0);
    gd.nmain = (uint16_t)nmain;
    gd.ncond = (uint16_t)cond.size();
    bgroups.push_back(gd);
  }
  flush();
}

bool dynamic_tiles();

// one run: the gates `run` (indices into `in`, execution order) all have their target in the tile `td`.
// One launch (one sweep) per batch of descriptors; the descriptors travel as a kernel parameter.
template <typename T, int kThreads>
int launch_run_shape(iqsb_state *st, const iqsb_fgate *in, const std::vector<int> &run, const TileDesc &td, bool reorder) {
  iqsb_ctx *ctx = st->ctx;
  std::vector<unsigned char> blob;
  int nbatches = 0;
  build_batches<T>(in, run, td, reorder, blob, nbatches);
  if (nbatches == 0) return IQSB_OK;
  const size_t smem = (size_t)sizeof(Cx<T>) << td.nS;
  auto kernel = ctx->arith == IQSB_ARITH_FMA ? k_fused<T, true, kThreads> : k_fused<T, false, kThreads>;
  IQSB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 1;
  IQSB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, smem));
  if (per_sm < 1) per_sm = 1;
  const uint64_t nouter = st->local_amps >> td.nS;
  const uint64_t cap = (uint64_t)ctx->num_sms * per_sm;
  const unsigned grid = (unsigned)(nouter < cap ? nouter : cap);
  const char *dbg = getenv("IQS_B200_FUSED_DEBUG");
  const int no_io = dbg && strcmp(dbg, "noio") == 0;
  static thread_local RunParams<T> params;  // host copy; the launch copies it into the parameter bank
  for (int b = 0; b < nbatches; ++b) {
    const unsigned char *base = blob.data() + (size_t)b * batch_stride<T>();
    const BatchHdr *h = reinterpret_cast<const BatchHdr *>(base);
    const GroupDesc *gd = reinterpret_cast<const GroupDesc *>(base + sizeof(BatchHdr));
    const FGate<T> *fg = reinterpret_cast<const FGate<T> *>(base + sizeof(BatchHdr) + kBatchGroups * sizeof(GroupDesc));
    memset(&params, 0, sizeof(params));
    params.ngroups = h->ngroups;
    for (int gi = 0; gi < h->ngroups; ++gi) {
      GroupHdr &o = params.hdr[gi];
      memcpy(o.p, gd[gi].p, sizeof(o.p));
      memcpy(o.w, gd[gi].w, sizeof(o.w));
      o.gate_first = gd[gi].gate_first;
      o.nmain = gd[gi].nmain;
      o.ncond = gd[gi].ncond;
      o.log2_threads = gd[gi].log2_threads;
      o.c0 = gd[gi].c0;
      memcpy(&params.lohi[gi][0], gd[gi].lo, sizeof(gd[gi].lo));
      memcpy(&params.lohi[gi][32], gd[gi].hi, sizeof(gd[gi].hi));
    }
    memcpy(params.gates, fg, (size_t)h->ngates * sizeof(FGate<T>));
    unsigned long long *counter = nullptr;
    if (dynamic_tiles()) {
      if (!ctx->d_tile_counter) IQSB_CUDA(cudaMalloc((void **)&ctx->d_tile_counter, sizeof(unsigned long long)));
      counter = ctx->d_tile_counter;
      IQSB_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned long long), ctx->stream));
    }
    kernel<<<grid, kThreads, smem, ctx->stream>>>((Chunk<T> *)st->d, nouter, td, counter, no_io, params);
    IQSB_TRY(iqsb_check_launch(ctx, "k_fused", 2.0 * (double)st->local_amps * st->amp_bytes()));
  }
  return IQSB_OK;
}

// tiles of 2^12 amplitudes: 256 threads, 3 CTAs per SM; smaller tiles: 128 threads, 6 CTAs per SM
template <typename T>
int launch_run(iqsb_state *st, const iqsb_fgate *in, const std::vector<int> &run, const TileDesc &td, bool reorder) {
  return td.nS > kSmallTile ? launch_run_shape<T, 256>(st, in, run, td, reorder) : launch_run_shape<T, 128>(st, in, run, td, reorder);
}

// The planner.  order = gate indices in execution order, run_end[r] = one past the last entry of
// run r in `order`, tiles[r*16] = number of tile positions, tiles[r*16 + 1 ..] = the positions.
int plan_runs_k(const iqsb_fgate *gates, int ngates, unsigned log2_local, bool reorder, unsigned tile_log2, int *order, int *run_end, uint8_t *tiles,
                int max_runs, int *nruns) {
  const unsigned K = log2_local < tile_log2 ? log2_local : tile_log2;
  const unsigned low = log2_local < (unsigned)kLow ? log2_local : (unsigned)kLow;
  std::vector<char> done((size_t)ngates, 0);
  int r = 0, first_pending = 0, nordered = 0;
  while (nordered < ngates) {
    IQSB_REQUIRE(r < max_runs, "iqsb_plan_fused: more than %d runs", max_runs);
    while (done[first_pending]) ++first_pending;
    bool in[64] = {false};
    unsigned cnt = 0;
    for (unsigned b = 0; b < low; ++b) { in[b] = true; ++cnt; }
    const int run_first = nordered;
    uint64_t skipped_perm_q = 0, skipped_arith_q = 0;
    bool arith_skipped = false;
    int looked = 0;
    for (int i = first_pending; i < ngates; ++i) {
      if (done[i]) continue;
      const iqsb_fgate &g = gates[i];
      const unsigned t = (unsigned)g.target;
      IQSB_REQUIRE(t < log2_local, "iqsb_fused: gate %d: target %u is not a local position", i, t);
      const uint64_t qmask = (1ull << t) | (g.kind == 1 ? 1ull << (unsigned)g.control : 0ull);
      const bool perm = is_xexact(g.m);
      // pulled across what was skipped only if that is exact: a permutation commutes without
      // rounding with gates on other qubits; two arithmetic gates never swap
      bool ok;
      if (perm) ok = !(qmask & (skipped_perm_q | skipped_arith_q));
      else ok = !arith_skipped && !(qmask & skipped_perm_q);
      const bool fits = in[t] || cnt < K;
      if (ok && fits) {
        if (!in[t]) { in[t] = true; ++cnt; }
        order[nordered++] = i;
        done[i] = 1;
        continue;
      }
      if (!reorder) break;
      if (perm) skipped_perm_q |= qmask;
      else { skipped_arith_q |= qmask; arith_skipped = true; }
      if (++looked > kReorderWindow) break;
    }
    // use the spare slots for controls of the run (cheaper inside the tile), then for low positions
    for (int k = run_first; k < nordered && cnt < K; ++k) {
      const iqsb_fgate &g = gates[order[k]];
      if (g.kind == 1 && (unsigned)g.control < log2_local && !in[g.control]) { in[g.control] = true; ++cnt; }
    }
    for (unsigned b = 0; b < log2_local && cnt < K; ++b)
      if (!in[b]) { in[b] = true; ++cnt; }
    uint8_t *td = tiles + r * 16;
    td[0] = (uint8_t)cnt;
    int n = 0;
    for (unsigned b = 0; b < log2_local; ++b)
      if (in[b]) td[1 + n++] = (uint8_t)b;
    for (; n < 15; ++n) td[1 + n] = 0;
    run_end[r++] = nordered;
  }
  *nruns = r;
  return IQSB_OK;
}

// The tile exponent a batch of gates is run with: 2^11 (the faster CTA shape) unless that costs a run more.
int plan_runs(const iqsb_fgate *gates, int ngates, unsigned log2_local, bool reorder, int *order, int *run_end, uint8_t *tiles, int max_runs, int *nruns) {
  if (ngates == 0 || log2_local <= (unsigned)kSmallTile || kTile <= kSmallTile)
    return plan_runs_k(gates, ngates, log2_local, reorder, (unsigned)kTile, order, run_end, tiles, max_runs, nruns);
  std::vector<int> order_s((size_t)ngates), end_s((size_t)ngates);
  std::vector<uint8_t> tiles_s((size_t)ngates * 16);
  int nruns_s = 0;
  IQSB_TRY(plan_runs_k(gates, ngates, log2_local, reorder, (unsigned)kSmallTile, order_s.data(), end_s.data(), tiles_s.data(), ngates, &nruns_s));
  IQSB_TRY(plan_runs_k(gates, ngates, log2_local, reorder, (unsigned)kTile, order, run_end, tiles, max_runs, nruns));
  if (nruns_s <= *nruns) {
    IQSB_REQUIRE(nruns_s <= max_runs, "iqsb_plan_fused: more than %d runs", max_runs);
    memcpy(order, order_s.data(), sizeof(int) * (size_t)ngates);
    memcpy(run_end, end_s.data(), sizeof(int) * (size_t)nruns_s);
    memcpy(tiles, tiles_s.data(), (size_t)nruns_s * 16);
    *nruns = nruns_s;
  }
  return IQSB_OK;
}

bool dynamic_tiles() {  // IQS_B200_FUSED_DYNAMIC=0: static round-robin of tiles over CTAs
  const char *e = getenv("IQS_B200_FUSED_DYNAMIC");
  return !(e && *e == '0');
}

bool reorder_default() {
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("IQS_B200_FUSED_REORDER");
    v = (e && *e == '0') ? 0 : 1;
  }
  return v != 0;
}

}  // namespace

extern "C" int iqsb_fused_max_log2tile(const iqsb_state *st) {
  if (!st) return 0;
  int k = kTile;
  if ((unsigned)k > st->log2_local) k = (int)st->log2_local;
  return k;
}

// Pure host function: the in-order cut of a batch into runs (no gate changes place).
extern "C" int iqsb_plan_fused(const iqsb_fgate *gates, int ngates, unsigned log2_local, int *run_end, uint8_t *tiles, int max_runs, int *nruns) {
  IQSB_REQUIRE((gates || ngates == 0) && run_end && tiles && nruns && max_runs > 0, "iqsb_plan_fused: null argument");
  std::vector<int> order((size_t)(ngates > 0 ? ngates : 1));
  return plan_runs(gates, ngates, log2_local, false, order.data(), run_end, tiles, max_runs, nruns);
}

// Pure host function: the plan iqsb_fused executes -- exact X / CNOT gates may move ahead of gates
// on other qubits (reorder != 0); order[] lists the gate indices in execution order.
extern "C" int iqsb_plan_fused_order(const iqsb_fgate *gates, int ngates, unsigned log2_local, int reorder, int *order, int *run_end, uint8_t *tiles,
                                     int max_runs, int *nruns) {
  IQSB_REQUIRE((gates || ngates == 0) && order && run_end && tiles && nruns && max_runs > 0, "iqsb_plan_fused_order: null argument");
  for (int i = 0; i < ngates; ++i)
    IQSB_REQUIRE(gates[i].target >= 0 && gates[i].target < 64 && (gates[i].kind != 1 || (gates[i].control >= 0 && gates[i].control < 64)),
                 "iqsb_plan_fused_order: gate %d has a bad position", i);
  return plan_runs(gates, ngates, log2_local, reorder != 0, order, run_end, tiles, max_runs, nruns);
}

// Pure host function: the complete schedule iqsb_fused builds for a ComplexDP register -- runs, groups,
// the arithmetic class and the kind of control of every gate -- decoded from the very descriptors the
// kernel reads.  out[k] describes the k-th gate executed; group_pos[4 g + j] = the position held by
// register bit j of group g (groups numbered over all runs).
extern "C" int iqsb_plan_fused_trace(const iqsb_fgate *gates, int ngates, unsigned log2_local, int reorder, iqsb_fused_trace *out, uint8_t *group_pos,
                                     int *ngroups) {
  IQSB_REQUIRE((gates || ngates == 0) && out && group_pos && ngroups, "iqsb_plan_fused_trace: null argument");
  IQSB_REQUIRE(log2_local >= (unsigned)kRegBits + 1, "iqsb_plan_fused_trace: shards below 2^%d amplitudes are not tiled", kRegBits + 1);
  *ngroups = 0;
  if (ngates == 0) return IQSB_OK;
  for (int i = 0; i < ngates; ++i)
    IQSB_REQUIRE((gates[i].kind == 0 || gates[i].kind == 1) && gates[i].target >= 0 && (unsigned)gates[i].target < log2_local &&
                     (gates[i].kind == 0 || (gates[i].control >= 0 && (unsigned)gates[i].control < log2_local && gates[i].control != gates[i].target)),
                 "iqsb_plan_fused_trace: gate %d is not a gate on local positions", i);
  std::vector<int> order((size_t)ngates), run_end((size_t)ngates);
  std::vector<uint8_t> tiles((size_t)ngates * 16);
  int nruns = 0;
  IQSB_TRY(plan_runs(gates, ngates, log2_local, reorder != 0, order.data(), run_end.data(), tiles.data(), ngates, &nruns));
  int first = 0, k = 0, g = 0;
  for (int r = 0; r < nruns; ++r) {
    TileDesc td;
    td.nS = tiles[r * 16];
    for (int j = 0; j < kTile; ++j) td.pos[j] = tiles[r * 16 + 1 + j];
    std::vector<int> run(order.begin() + first, order.begin() + run_end[r]);
    first = run_end[r];
    std::vector<unsigned char> blob;
    int nbatches = 0;
    build_batches<double>(gates, run, td, reorder != 0, blob, nbatches);
    for (int b = 0; b < nbatches; ++b) {
      const unsigned char *base = blob.data() + (size_t)b * batch_stride<double>();
      const BatchHdr *h = reinterpret_cast<const BatchHdr *>(base);
      const GroupDesc *gd = reinterpret_cast<const GroupDesc *>(base + sizeof(BatchHdr));
      const FGate<double> *fg = reinterpret_cast<const FGate<double> *>(base + sizeof(BatchHdr) + kBatchGroups * sizeof(GroupDesc));
      for (int gi = 0; gi < h->ngroups; ++gi, ++g) {
        for (int j = 0; j < 4; ++j) group_pos[4 * g + j] = 255;
        for (int j = 0; j < kRegBits; ++j) {
          const unsigned one_hot = swz(gd[gi].p[j]);  // the swizzle is an involution
          group_pos[4 * g + j] = td.pos[__builtin_ctz(one_hot)];
        }
        for (int q = gd[gi].gate_first; q < gd[gi].gate_first + gd[gi].gate_count; ++q, ++k) {
          IQSB_REQUIRE(k < ngates, "iqsb_plan_fused_trace: internal error (more gates scheduled than given)");
          out[k].gate = (int32_t)fg[q].origin;
          out[k].run = r;
          out[k].group = g;
          out[k].cls = fg[q].cls;
          out[k].tbit = fg[q].tbit;
          out[k].ckind = fg[q].ckind;
          out[k].c = fg[q].c;
          out[k].trail = fg[q].trail;
          out[k].pad[0] = out[k].pad[1] = out[k].pad[2] = 0;
        }
      }
    }
  }
  IQSB_REQUIRE(k == ngates, "iqsb_plan_fused_trace: internal error (%d of %d gates scheduled)", k, ngates);
  *ngroups = g;
  return IQSB_OK;
}

// Pure host function: the raw descriptors of the schedule (what the kernel is given), for tools and
// for the CPU model of the kernel in tests/fused_model.py.  Layout of out[]: int32 nruns; per run: int32 nS,
// uint8 pos[12], int32 nbatches, then nbatches blocks of {int32 ngroups, ngates, 0, 0; GroupDesc[24];
// FGate<double>[48]} exactly as built for the launch.  *used = bytes written (or needed when cap is too small).
extern "C" int iqsb_plan_fused_dump(const iqsb_fgate *gates, int ngates, unsigned log2_local, int reorder, void *out, size_t cap, size_t *used) {
  IQSB_REQUIRE((gates || ngates == 0) && used, "iqsb_plan_fused_dump: null argument");
  IQSB_REQUIRE(log2_local >= (unsigned)kRegBits + 1, "iqsb_plan_fused_dump: shards below 2^%d amplitudes are not tiled", kRegBits + 1);
  for (int i = 0; i < ngates; ++i)
    IQSB_REQUIRE((gates[i].kind == 0 || gates[i].kind == 1) && gates[i].target >= 0 && (unsigned)gates[i].target < log2_local &&
                     (gates[i].kind == 0 || (gates[i].control >= 0 && (unsigned)gates[i].control < log2_local && gates[i].control != gates[i].target)),
                 "iqsb_plan_fused_dump: gate %d is not a gate on local positions", i);
  std::vector<unsigned char> buf;
  auto put = [&](const void *p, size_t n) { buf.insert(buf.end(), (const unsigned char *)p, (const unsigned char *)p + n); };
  int nruns = 0;
  std::vector<int> order((size_t)(ngates > 0 ? ngates : 1)), run_end((size_t)(ngates > 0 ? ngates : 1));
  std::vector<uint8_t> tiles((size_t)(ngates > 0 ? ngates : 1) * 16);
  if (ngates > 0) IQSB_TRY(plan_runs(gates, ngates, log2_local, reorder != 0, order.data(), run_end.data(), tiles.data(), ngates, &nruns));
  int32_t v = nruns;
  put(&v, 4);
  int first = 0;
  for (int r = 0; r < nruns; ++r) {
    TileDesc td;
    td.nS = tiles[r * 16];
    for (int j = 0; j < kTile; ++j) td.pos[j] = tiles[r * 16 + 1 + j];
    std::vector<int> run(order.begin() + first, order.begin() + run_end[r]);
    first = run_end[r];
    std::vector<unsigned char> blob;
    int nbatches = 0;
    build_batches<double>(gates, run, td, reorder != 0, blob, nbatches);
    v = td.nS;
    put(&v, 4);
    uint8_t pos12[12] = {0};
    for (int j = 0; j < kTile && j < 12; ++j) pos12[j] = td.pos[j];
    put(pos12, 12);
    v = nbatches;
    put(&v, 4);
    put(blob.data(), blob.size());
  }
  *used = buf.size();
  if (out && cap >= buf.size()) memcpy(out, buf.data(), buf.size());
  else IQSB_REQUIRE(out == nullptr, "iqsb_plan_fused_dump: buffer of %zu bytes is too small (%zu needed)", cap, buf.size());
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
  if (st->log2_local < (unsigned)kRegBits + 1) {  // nothing to tile
    for (int i = 0; i < ngates; ++i) {
      if (gates[i].kind == 0) IQSB_TRY(iqsb_gate1(st, (unsigned)gates[i].target, gates[i].m, 0, st->local_amps));
      else IQSB_TRY(iqsb_cgate1(st, (unsigned)gates[i].control, (unsigned)gates[i].target, gates[i].m, 0, st->local_amps));
    }
    return IQSB_OK;
  }
  std::vector<int> order((size_t)ngates), run_end((size_t)ngates);
  std::vector<uint8_t> tiles((size_t)ngates * 16);
  int nruns = 0;
  const bool reorder = reorder_default();
  IQSB_TRY(plan_runs(gates, ngates, st->log2_local, reorder, order.data(), run_end.data(), tiles.data(), ngates, &nruns));
  int first = 0;
  for (int r = 0; r < nruns; ++r) {
    TileDesc td;
    td.nS = tiles[r * 16];
    for (int k = 0; k < kTile; ++k) td.pos[k] = tiles[r * 16 + 1 + k];
    std::vector<int> run(order.begin() + first, order.begin() + run_end[r]);
    int rc = st->dtype == IQSB_F64 ? launch_run<double>(st, gates, run, td, reorder) : launch_run<float>(st, gates, run, td, reorder);
    if (rc != IQSB_OK) return rc;
    first = run_end[r];
  }
  return IQSB_OK;
}
