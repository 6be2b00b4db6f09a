/*
 * iqsb.h -- C ABI of libiqs_b200.so: the B200 (sm_100a) state-vector engine that sits
 * under the Intel-QS `iqs::QubitRegister<Type>` class.
 *
 * Nothing like this exists in the reference: its seam is a set of C++ function
 * templates (Loop_SN/Loop_DN/Loop_TN/ScaleState, reference include/highperfkernels.hpp:10-34)
 * plus loops written inline in the QubitRegister methods.  Each entry point below names
 * the reference interface it replaces (file:line are relative to the reference tree).
 *
 * Conventions
 *  - every function returns 0 on success and a negative code on failure; the message of
 *    the last failure of the calling thread is available from iqsb_last_error().
 *  - amplitudes are interleaved (re, im); a register shard holds `local_amps` amplitudes of
 *    `double` (IQSB_F64, ComplexDP) or `float` (IQSB_F32, ComplexSP).
 *  - 2x2 matrices are passed as `const double m[8]`, row-major, (re, im) per entry:
 *    m = {m00.re, m00.im, m01.re, m01.im, m10.re, m10.im, m11.re, m11.im}.
 *    For IQSB_F32 registers the entries are rounded to float once, on the host.
 *  - `pos` arguments are *positions* (data-qubit indices, reference qureg.hpp:80-85) and
 *    are local: pos < log2(local_amps), unless the function name says `_global`.
 *  - all work is enqueued on the context's stream; functions returning a scalar, and
 *    iqsb_download / iqsb_sync, synchronise that stream.
 *  - there is no CPU fallback: without a CUDA device iqsb_init fails.
 */
#ifndef IQSB_H
#define IQSB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct iqsb_ctx iqsb_ctx;     /* one per process: device, stream, NCCL communicator */
typedef struct iqsb_state iqsb_state; /* one register shard (+ optional tmp area)            */

enum { IQSB_F64 = 0, IQSB_F32 = 1 };
enum {
  IQSB_MEM_DEVICE = 0, /* cudaMalloc: fastest, IPC-exportable (needed for nranks > 1)  */
  IQSB_MEM_MANAGED = 1 /* cudaMallocManaged, device-preferred: host pointer is valid   */
};
enum { IQSB_OK = 0, IQSB_ERR_CUDA = -1, IQSB_ERR_NCCL = -2, IQSB_ERR_ARG = -3, IQSB_ERR_STATE = -4, IQSB_ERR_PEER = -5 };
enum { IQSB_SUM = 0, IQSB_MAX = 1 };
#define IQSB_UNIQUE_ID_BYTES 128

/* ---- library ------------------------------------------------------------------------ */
int iqsb_version(void);
const char *iqsb_last_error(void);

/* ---- context: replaces iqs::mpi::Environment Init/Finalize/rank queries
 *      (include/mpi_env.hpp:46-145, src/mpi_env.cpp:95-230) ----------------------------- */
/* rank 0 creates the NCCL id; the launcher hands the 128 bytes to the other ranks. */
int iqsb_unique_id(void *out_128_bytes);
/* device < 0 means "rank % visible devices". uid may be NULL when nranks == 1. */
int iqsb_init(int rank, int nranks, const void *uid, int device, iqsb_ctx **out);
int iqsb_finalize(iqsb_ctx *ctx);
int iqsb_rank(const iqsb_ctx *ctx);
int iqsb_nranks(const iqsb_ctx *ctx);
int iqsb_device(const iqsb_ctx *ctx);
int iqsb_sync(iqsb_ctx *ctx);
/* free and total HBM of the context's device, in bytes */
int iqsb_mem_info(iqsb_ctx *ctx, uint64_t *free_bytes, uint64_t *total_bytes);
/* adopt an external cudaStream_t (e.g. torch's current stream) so that the caller's
 * CUDA events bracket our launches; NULL restores the context's own stream. */
int iqsb_set_stream(iqsb_ctx *ctx, void *cuda_stream);
void *iqsb_get_stream(iqsb_ctx *ctx);
/* Arithmetic of the fused kernel.  IQSB_ARITH_EXACT (default): separately rounded products and
 * sums in the reference's operation order -- results bit-identical to the reference built for
 * baseline x86-64.  IQSB_ARITH_FMA: contracted multiply-adds, what the reference's IqsNative=ON
 * (-march=native) build lets its compiler do (CMakeLists.txt:205-215); results agree to ~1e-16 and
 * the in-tile phase of iqsb_fused issues 16 instead of 28 FP64 instructions per pair.  The initial
 * mode comes from the environment variable IQS_B200_ARITH ("exact" | "fma"). */
#define IQSB_ARITH_EXACT 0
#define IQSB_ARITH_FMA 1
int iqsb_set_arith(iqsb_ctx *ctx, int mode);
int iqsb_get_arith(const iqsb_ctx *ctx);
/* number of kernels this context launched since creation (bench.py: gpu_launches). */
uint64_t iqsb_launch_count(const iqsb_ctx *ctx);
/* Errors raised on the device since the context was created: IQSB_ERR_PEER when a rendezvous between
 * GPUs (the flag barrier that brackets every peer-memory kernel) gave up because a partner did not
 * arrive within IQS_B200_BARRIER_TIMEOUT_S seconds (default 300, 0 = wait for ever) -- the reference's
 * MPI job would hang in MPI_Sendrecv there.  Also returned by iqsb_sync, iqsb_barrier and every
 * reduction, which is where a program notices. */
int iqsb_check(iqsb_ctx *ctx);
/* Per-kernel-class device timing with CUDA events on the engine's stream (what bench.py's roofline
 * block is computed from; the reference's counterpart is the Timer of include/timer.hpp, filled by
 * QubitRegister::EnableStatistics).  iqsb_profile(ctx, 1) clears and starts, iqsb_profile(ctx, 0)
 * stops; iqsb_profile_read writes JSON text {"overflow": bool, "classes": [{"name", "launches",
 * "ms", "bytes" (algorithmic bytes, 0 if not stated)}]} into out[cap]. */
int iqsb_profile(iqsb_ctx *ctx, int on);
int iqsb_profile_read(iqsb_ctx *ctx, char *out, size_t cap);
/* device-side timing on the context's stream (CUDA events). */
int iqsb_timer_start(iqsb_ctx *ctx);
int iqsb_timer_stop(iqsb_ctx *ctx, double *elapsed_ms);
/* numbered CUDA events on the context's stream (slot < 4096): record now, read the elapsed time
 * between two recorded slots later (the call synchronises on the later event). */
int iqsb_event_record(iqsb_ctx *ctx, int slot);
int iqsb_event_elapsed(iqsb_ctx *ctx, int slot_from, int slot_to, double *elapsed_ms);

/* ---- scalar collectives: replace MPI_Allreduce_x / MPI_Bcast_x / MPI_Barrier
 *      (include/mpi_utils.hpp:34-78, src/mpi_env.cpp:499-509) ------------------------- */
int iqsb_allreduce_f64(iqsb_ctx *ctx, double *inout, int n, int op);
int iqsb_bcast_f64(iqsb_ctx *ctx, double *inout, int n, int root);
int iqsb_barrier(iqsb_ctx *ctx);

/* ---- memory: replaces QubitRegister::Allocate/Resize/dtor
 *      (src/qureg_init.cpp:44-75,148-187,449-457) ------------------------------------- */
int iqsb_alloc(iqsb_ctx *ctx, uint64_t local_amps, uint64_t tmp_amps, int dtype, int mem_kind,
               iqsb_state **out);
int iqsb_free(iqsb_state *st);
uint64_t iqsb_local_amps(const iqsb_state *st);
int iqsb_dtype(const iqsb_state *st);
void *iqsb_device_ptr(iqsb_state *st);
/* host-dereferenceable pointer to the shard (IQSB_MEM_MANAGED only, else NULL). */
void *iqsb_host_ptr(iqsb_state *st);
/* make the shard resident in HBM again after host accesses (managed memory only; no-op else). */
int iqsb_prefetch_device(iqsb_state *st);
int iqsb_upload(iqsb_state *st, const void *host_amps, uint64_t first_amp, uint64_t count);
int iqsb_download(iqsb_state *st, void *host_amps, uint64_t first_amp, uint64_t count);
/* dst[i] = src[i] (copy constructor, src/qureg_init.cpp:364-379). */
int iqsb_copy(iqsb_state *dst, const iqsb_state *src);

/* ---- initialisation: Initialize("base"/"++++") and friends
 *      (src/qureg_init.cpp:218-347; qureg_utils.cpp:173-184) -------------------------- */
int iqsb_fill_const(iqsb_state *st, double re, double im); /* InitializationWithSameAmplitudeEverywhere */
int iqsb_set_amp(iqsb_state *st, uint64_t local_index, double re, double im); /* SetGlobalAmplitude on the owner */
int iqsb_get_amp(iqsb_state *st, uint64_t local_index, double *re, double *im);
/* state[i] = U[-1,1) + i U[-1,1) from a counter-based generator (splitmix64 of seed and the
 * global amplitude index) -- benchmark/test input only, not the reference's mt19937 stream. */
int iqsb_fill_random(iqsb_state *st, uint64_t seed, uint64_t global_offset);

/* ---- gate kernels -------------------------------------------------------------------- */
/* 1-qubit gate on local position pos over the index range [sind, eind):
 * replaces Loop_DN(sind, eind, pos, state, state, 0, 1<<pos, m, ...)
 * (src/highperfkernels.cpp:287-380; called from qureg_apply1qubitgate.cpp:206). */
int iqsb_gate1(iqsb_state *st, unsigned pos, const double m[8], uint64_t sind, uint64_t eind);
/* controlled 1-qubit gate, both positions local: replaces the two Loop_TN calls of
 * ApplyControlled1QubitGate_helper (src/qureg_applyctrl1qubitgate.cpp:312-345,
 * src/highperfkernels.cpp:397-486). */
int iqsb_cgate1(iqsb_state *st, unsigned cpos, unsigned tpos, const double m[8], uint64_t sind,
                uint64_t eind);
/* swap-family gate: 2x2 `m` on the {pos1=1,pos2=0} <-> {pos1=0,pos2=1} subspace with
 * pos1 < pos2 both local: replaces Loop_TN in ApplySwap_helper (src/qureg_applyswap.cpp:203-206). */
int iqsb_swap2x2(iqsb_state *st, unsigned pos1, unsigned pos2, const double m[8]);
/* two-qubit diagonal gate, amp *= d[2*bit(pos1)+bit(pos2)] (src/qureg_applydiag.cpp:157-224).
 * A position >= log2(local_amps) is "global": its bit is taken from glb_start. */
int iqsb_diag2(iqsb_state *st, unsigned pos1, unsigned pos2, const double d[8], uint64_t glb_start);
/* state[i] *= s for i in [start, end): ScaleState (src/highperfkernels.cpp:501-520). */
int iqsb_scale(iqsb_state *st, const double s[2], uint64_t start, uint64_t end);
/* amp *= s for the amplitudes whose bit `pos` is 1 (diagonal 1-qubit gate diag(1, s) and, with
 * cpos >= 0, its controlled form); used for diag(d0,d1) gates as scale-by-bit. */
int iqsb_phase_by_bit(iqsb_state *st, int cpos, unsigned pos, const double d0[2], const double d1[2]);
/* general 4x4 gate on local positions (high, low): Apply2QubitGate (src/qureg_apply2qubitgate.cpp:15-73);
 * m is 16 complex numbers row-major, basis index t = 2*bit(pos_high) + bit(pos_low). */
int iqsb_gate2(iqsb_state *st, unsigned pos_high, unsigned pos_low, const double m[32]);

/* ---- gate fusion: replaces ApplyFusedGates (src/qureg_fusion.cpp:55-94) -------------- */
typedef struct iqsb_fgate {
  int32_t kind; /* 0 = 1-qubit gate on `target`; 1 = controlled gate (control, target) */
  int32_t control;
  int32_t target; /* target position (local) */
  int32_t pad;
  double m[8];
} iqsb_fgate;
/* Apply `ngates` gates in order with as few sweeps over HBM as possible.  Targets and controls may
 * be ANY local positions: the batch is cut into runs of gates whose targets fit in one
 * shared-memory tile (2^12 amplitudes = the 4 lowest positions + 8 positions chosen per run, or 2^11
 * with 7 when that costs no extra run: the smaller tile runs 6 CTAs per SM); each run costs one read
 * and one write of the state.  Inside a run, gates on up to three tile bits are
 * applied in registers per shared-memory round trip, with arithmetic specialised to the zero
 * structure of each matrix (values identical to the full evaluation for finite amplitudes). */
int iqsb_fused(iqsb_state *st, const iqsb_fgate *gates, int ngates);
/* largest tile exponent (12, or log2(local_amps) for tiny shards) */
int iqsb_fused_max_log2tile(const iqsb_state *st);
/* Pure host function: the runs iqsb_fused would execute.  run_end[r] = one past the last gate of
 * run r; tiles[16 r] = number of tile positions, tiles[16 r + 1 ..] = the positions (ascending); tiles holds 16 bytes per run. */
int iqsb_plan_fused(const iqsb_fgate *gates, int ngates, unsigned log2_local, int *run_end, uint8_t *tiles, int max_runs, int *nruns);
/* Pure host function: the plan iqsb_fused really executes.  Runs are cut in program order, except
 * that a pure-permutation gate (X / CNOT: matrix entries exactly 0 and 1) may move ahead of gates on
 * OTHER qubits into an earlier run -- that commutation involves no rounding, so the result is bit
 * for bit the one of the program order (reorder = 0 switches it off; IQS_B200_FUSED_REORDER=0 does
 * so for iqsb_fused).  order[k] = index of the k-th gate executed, run_end[r] = one past the last
 * entry of run r in order[], tiles as for iqsb_plan_fused. */
int iqsb_plan_fused_order(const iqsb_fgate *gates, int ngates, unsigned log2_local, int reorder, int *order, int *run_end, uint8_t *tiles,
                          int max_runs, int *nruns);

/* Pure host function: the complete schedule of iqsb_fused for a ComplexDP register, decoded from the
 * descriptors the kernel reads.  out[k] = the k-th gate executed: its index in gates[], its run, its
 * group (numbered over all runs), its arithmetic class (0 general, 1 real, 2 diagonal, 3 diag(1,d),
 * 4 anti-diagonal, 5 exact X, 6 real diagonal + imaginary off-diagonal, 7 sqrt X, 8 sqrt Y), the
 * register bit of its target, the kind of its control (0 none, 1 register bit, 2 thread bit,
 * 3 bit of the tile's base index) and whether it is applied to the registers or folded into the
 * write-back addresses of its group.  group_pos[4 g + j] = position held by register bit j of group g
 * (255 = unused); group_pos holds 4 bytes per gate at most. */
typedef struct iqsb_fused_trace {
  int32_t gate, run, group;
  uint8_t cls, tbit, ckind, c;
  uint8_t trail; /* 0: applied to the registers; 1 / 2: exact X / CNOT at the end of its group, folded into the
                  * write-back addresses (1: conditional offset, 2: absorbed into the basis) -- such gates take
                  * effect in program order whatever their place in this list */
  uint8_t pad[3];
} iqsb_fused_trace;
int iqsb_plan_fused_trace(const iqsb_fgate *gates, int ngates, unsigned log2_local, int reorder, iqsb_fused_trace *out, uint8_t *group_pos,
                          int *ngroups);

/* Pure host function: the raw descriptors of that schedule (tile positions, group headers with their slot
 * tables and write-back bases, gates with class / control / matrix), as the kernel receives them; the
 * layout is documented at the definition (csrc/kernels_fused.cu) and mirrored by tests/fused_model.py, a
 * CPU model of the kernel used to check the schedule against the oracle without a GPU.  out == NULL:
 * only *used (the size needed) is returned. */
int iqsb_plan_fused_dump(const iqsb_fgate *gates, int ngates, unsigned log2_local, int reorder, void *out, size_t cap, size_t *used);

/* ---- reductions (warp-shuffle + fixed-order second stage; deterministic run to run) -- */
/* sum |a|^2 over local amplitudes with bit pos == 1: GetProbability (src/qureg_measure.cpp:150-167) */
int iqsb_prob1(iqsb_state *st, unsigned pos, double *out);
/* sum_i (-1)^popcount((glb_start+i) & mask) |a_i|^2: ExpectationValue (src/qureg_expectval.cpp:173-185) */
int iqsb_parity_expect(iqsb_state *st, uint64_t mask, uint64_t glb_start, double *out);
/* All marginals in ONE read of the shard: out[0] = sum |a|^2, out[1 + q] = sum of |a_i|^2 over local
 * indices with bit q set, q < log2(local_amps); nout >= 1 + log2(local_amps).  n calls of
 * GetProbability (src/qureg_measure.cpp:135-178) cost one sweep instead of n. */
int iqsb_prob_all(iqsb_state *st, double *out, int nout);
/* Read-only expectation value of a Pauli string: X on the bits of xmask, Y on ymask, Z on zmask
 * (masks over the GLOBAL index; X / Y bits must be local positions, Z bits may be rank bits, taken
 * from glb_start).  Replaces the basis-change sweeps of ExpectationValue
 * (src/qureg_expectval.cpp:148-210: 2k gate sweeps + 1 read) by one read; the state is not touched.
 * out[0] = the expectation value (local part), out[1] = sum |a|^2 of the shard from the same read:
 * the reference's 1-qubit wrappers return 1 - 2 P(1) = <P> + (1 - norm^2) (src/qureg_expectval.cpp:18-66). */
int iqsb_pauli_expect(iqsb_state *st, uint64_t xmask, uint64_t ymask, uint64_t zmask, uint64_t glb_start, double out[2]);
/* sum |a|^2: ComputeNorm before sqrt (src/qureg_utils.cpp:236-255) */
int iqsb_norm2(iqsb_state *st, double *out);
/* sum conj(b_i) a_i: ComputeOverlap (src/qureg_utils.cpp:259-300); out = {re, im} */
int iqsb_overlap(iqsb_state *a, iqsb_state *b, double out[2]);
/* max_i |a_i - s b_i|: MaxAbsDiff (src/qureg_utils.cpp:35-72) */
int iqsb_maxabsdiff(iqsb_state *a, iqsb_state *b, const double s[2], double *out);
/* sum_i |a_i - b_i|^2: MaxL2NormDiff (src/qureg_utils.cpp:127-158) */
int iqsb_l2diff(iqsb_state *a, iqsb_state *b, double *out);
/* out[0] = any |a_i|^2 > tol with bit pos == 0, out[1] = same with bit pos == 1:
 * IsClassicalBit / GetClassicalValue (src/qureg_measure.cpp:19-81,183-262). pos >= log2(local)
 * is global: every amplitude counts for the bit value found in glb_start. */
int iqsb_any_above(iqsb_state *st, unsigned pos, double tol, uint64_t glb_start, int out[2]);
/* 1 if the two shards hold equal values (operator==, src/qureg_utils.cpp:17-31) */
int iqsb_equal(iqsb_state *a, iqsb_state *b, int *out);
/* -sum p ln p and the Google moments (src/qureg_utils.cpp:305-450): out[0] = sum -p ln p,
 * out[1] = sum -ln p, out[2..10] = sum p^k, k = 2..10 (local, unscaled). */
int iqsb_entropy_stats(iqsb_state *st, double out[11]);

/* ---- measurement / element-wise ------------------------------------------------------ */
/* zero the amplitudes whose bit pos != value: CollapseQubit (src/qureg_measure.cpp:92-126) */
int iqsb_collapse(iqsb_state *st, unsigned pos, int value);
/* a[i] += f * b[i]: AmplitudeWiseSum (src/qureg_utils.cpp:199-226) */
int iqsb_axpy(iqsb_state *a, const iqsb_state *b, const double f[2]);

/* ---- QAOA helpers (src/qaoa_features.cpp): a classical cost function lives in Re(diag[i]) -------- */
/* diag[i] = cut value of the bit string of global index glb_start + i (program order via
 * pos_of_qubit[q] = data position of program qubit q); adjacency is nverts x nverts, row-major.
 * weighted == 0: integer arithmetic of InitializeVectorAsMaxCutCostFunction (:58-118);
 * weighted != 0: the floating-point loop of InitializeVectorAsWeightedMaxCutCostFunction (:121-203),
 * same operation order.  Returns the largest local cut. */
int iqsb_qaoa_maxcut(iqsb_state *diag, unsigned nverts, const double *adjacency, int weighted, const uint8_t *pos_of_qubit,
                     uint64_t glb_start, double *max_cut_local);
/* psi[i] *= exp(-i gamma Re diag[i]): ImplementQaoaLayerBasedOnCostFunction (:255-266) */
int iqsb_qaoa_layer(iqsb_state *psi, const iqsb_state *diag, double gamma);
/* out[0] = sum Re diag |psi|^2, out[1] = sum (Re diag)^2 |psi|^2 (local): GetExpectationValue[Squared]FromCostFunction (:277-335) */
int iqsb_qaoa_expect(iqsb_state *psi, const iqsb_state *diag, double out[2]);
/* out[floor(Re diag / bin_width + eps)] += |psi|^2 (local): the three GetHistogramFromCostFunction* (:345-523) */
int iqsb_qaoa_histogram(iqsb_state *psi, const iqsb_state *diag, int nbins, double bin_width, double eps, double *out);

/* ---- qubit reordering: PermuteLocalQubits (src/qureg_permute.cpp:55-104) ------------- */
/* new[j] = old[i] where bit b of i becomes bit dst_bit[b] of j, b < log2(local_amps). */
int iqsb_permute_local(iqsb_state *st, const uint8_t *dst_bit, unsigned nbits);
/* Pure host function: the in-place tile phases iqsb_permute_local runs for this permutation.  Phase p:
 * out[25p] = number of tile positions nS, out[25p+1..] = the positions (ascending), out[25p+13..] = for
 * tile-local bit k the tile-local bit it moves to.  Used by the CPU tests of the planner. */
int iqsb_plan_permute(const uint8_t *dst_bit, unsigned nbits, uint8_t *out, int max_phases, int *nphases);

/* ---- distributed (nranks > 1): fused compute + NVLink peer access -------------------- */
/* Pure host function (no GPU needed): which pairs of a gate on a global qubit THIS rank updates.
 * The rank updates the pairs (s0[k + extra0], s1[k + extra1]) for every local index k whose bits
 * pos[i] equal val[i] (i < nfix); s0 is the shard of the rank on the "0 side" of the pair, s1 the
 * other; role says which of the two is this rank's own shard (0: s0 is mine, 1: s1 is mine).
 * kind 0: 1-qubit gate on global pos2; 1: controlled gate, control pos1 local, target pos2 global;
 * 2: swap-family gate on pos1 < pos2 with pos2 global. */
typedef struct iqsb_plan {
  int32_t active;  /* 0: this rank owns no pair of this gate (it still joins the rendezvous) */
  int32_t partner; /* rank whose shard is read and written over NVLink */
  int32_t role;
  int32_t nfix;
  uint32_t pos[3], val[3];
  uint64_t extra0, extra1;
  uint64_t npairs;
  uint64_t link_amps; /* amplitudes crossing the link per direction (algorithmic) */
} iqsb_plan;
int iqsb_plan_global(int kind, int rank, int nranks, unsigned M, unsigned pos1, unsigned pos2, iqsb_plan *out);
/* publish the shard to the peers (cudaIpc) -- collective over all ranks of the context. */
int iqsb_share(iqsb_state *st);
/* 1-qubit gate on global position pos >= M: replaces HP_Distrpair(P) (src/qureg_apply1qubitgate.cpp:18-169) */
int iqsb_gate1_global(iqsb_state *st, unsigned M, unsigned pos, const double m[8]);
/* controlled gate, control local (cpos < M), target global: replaces HP_Distrpair(C,T)
 * (src/qureg_applyctrl1qubitgate.cpp:24-220) */
int iqsb_cgate1_global(iqsb_state *st, unsigned M, unsigned cpos, unsigned tpos, const double m[8]);
/* swap-family gate with pos1 < pos2, pos2 >= M: replaces HP_DistrSwap (src/qureg_applyswap.cpp:247-480) */
int iqsb_swap2x2_global(iqsb_state *st, unsigned M, unsigned pos1, unsigned pos2, const double m[8]);
/* Exchange the contents of k (1..3) local positions with k global positions in ONE pass: afterwards
 * the qubit that lived at local position lpos[j] lives at global position gpos[j] and vice versa.
 * Pure data movement (bit-exact).  Generalises HP_DistrSwap with m = X (src/qureg_applyswap.cpp:247-480)
 * and replaces the pair-by-pair loop of PermuteByLocalGlobalExchangeOfQubitPairs
 * (src/qureg_permute.cpp:191-229).  NVLink traffic per rank and direction: (1 - 2^-k) * 16 B * local_amps.
 * Needs M >= k + 1.  This is what the host library's placement layer uses to bring a global qubit
 * in once and keep it local for the gates that follow. */
int iqsb_exchange_bits(iqsb_state *st, unsigned M, int k, const unsigned *lpos, const unsigned *gpos);
/* Pure host function (no GPU needed): what THIS rank moves in iqsb_exchange_bits.  For partner p the
 * rank trades its amplitudes whose local bits lpos[] spell mine[p] against the partner's amplitudes
 * whose bits spell theirs[p]; of each such pair of blocks it moves the half whose local bit
 * split_bit equals split_val[p] (the partner moves the other half). */
typedef struct iqsb_xplan {
  int32_t npartners; /* 2^k - 1 */
  int32_t split_bit;
  int32_t partner[7];
  int32_t split_val[7];
  uint64_t mine[7], theirs[7];
  uint64_t amps_per_partner; /* amplitudes this rank loads from (and stores to) each partner */
  uint64_t link_amps;        /* amplitudes crossing the link per direction (algorithmic) */
} iqsb_xplan;
int iqsb_plan_exchange(int rank, int nranks, unsigned M, int k, const unsigned *lpos, const unsigned *gpos, iqsb_xplan *out);
/* Placement planning, a pure host function (no GPU needed).  The host library keeps, next to the
 * reference's qubit_permutation (qubit -> position), a physical placement place[position] = bit of the
 * distributed index that currently holds it (bits >= M are rank bits).  Given the upcoming gates in
 * program order (positions; `diagonal` != 0 for diagonal matrices) this picks up to 3 positions to
 * bring in from rank bits and the local positions to evict for them -- the arguments of ONE
 * iqsb_exchange_bits call -- by Belady's rule on the next non-diagonal-target use.  Positions in
 * protect_mask must end up local (they are brought in first) and are never evicted.  last_use (may be
 * NULL) breaks ties towards the least recently used position; only positions held by local bits
 * >= min_evict_bit are evicted (low bits would give short runs on the link).
 * Automates what the reference leaves to the user: PermuteQubits / EmulateSwap
 * (src/qureg_permute.cpp:10-52, examples/communication_reduction_via_qubit_reordering.cpp:87-132). */
typedef struct iqsb_pgate {
  int32_t kind; /* 0: 1-qubit gate on target, 1: controlled gate */
  int32_t control;
  int32_t target;
  int32_t diagonal;
} iqsb_pgate;
int iqsb_plan_placement(const uint8_t *place, unsigned n, unsigned M, const iqsb_pgate *gates, int ngates, uint64_t protect_mask,
                        const uint64_t *last_use, unsigned min_evict_bit, unsigned *evict_pos, unsigned *bring_pos, int *k);
/* take part in a global-qubit step without owning pairs (ranks whose global control bit is 0,
 * src/qureg_applyctrl1qubitgate.cpp:359-380): same two rendezvous as the gate itself. */
int iqsb_idle_global(iqsb_state *st);
/* whole-shard move: this rank's shard goes to `dst_rank`, it receives `src_rank`'s:
 * replaces the Sendrecv loop of PermuteGlobalQubits (src/qureg_permute.cpp:174-185) */
int iqsb_permute_global(iqsb_state *st, int src_rank, int dst_rank);
/* the same move described by the permutation of the rank bits (content of rank bit b goes to rank bit
 * dst_rank_bit[b], nbits = log2(nranks)): every rank derives its source, its destination and the
 * path all ranks take from the table -- no collective is needed to agree. */
int iqsb_permute_global_bits(iqsb_state *st, const uint8_t *dst_rank_bit, unsigned nbits);
/* Pure host function: the plan of iqsb_permute_global_bits for one rank -- the rank it pulls its new
 * shard from, the rank its shard goes to, whether all ranks move in pairs, whether nothing moves. */
int iqsb_plan_permute_global_bits(int rank, int nranks, const uint8_t *dst_rank_bit, unsigned nbits, int *source, int *destination, int *pairwise,
                                  int *identity);
/* bytes this context moved over NVLink (peer loads + peer stores) since creation. */
uint64_t iqsb_nvlink_bytes(const iqsb_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* IQSB_H */
