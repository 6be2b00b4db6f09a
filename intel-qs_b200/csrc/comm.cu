// comm.cu -- the distributed half of the engine: one process per GPU, shards published to the
// peers with cudaIpc, gates on global qubits executed by ONE kernel per rank that reads and
// writes the partner's shard directly over NVLink (no tmp buffer, no send/recv phases).
//
// Reference code replaced:
//   HP_Distrpair(P)      src/qureg_apply1qubitgate.cpp:18-169   (4 MPI_Sendrecv + Loop_SN per chunk)
//   HP_Distrpair(C,T)    src/qureg_applyctrl1qubitgate.cpp:24-220
//   HP_DistrSwap         src/qureg_applyswap.cpp:247-480
//   PermuteGlobalQubits  src/qureg_permute.cpp:149-185
//   MPI_Allreduce_x / MPI_Bcast_x / barriers  include/mpi_utils.hpp:34-78, src/mpi_env.cpp:499-509
//
// Work split (same as the reference's i-task / j-task split, but without the exchange phases):
// a pair (a_k on the rank whose target bit is 0, b_k on the rank whose bit is 1) is owned by
// exactly one of the two ranks, chosen by one local "split" bit of k, so every pair is updated
// in place by one thread: that thread loads one local and one remote amplitude chunk and
// stores both.  Per rank and per direction the link carries 16 B x (pairs/2) of loads and the
// same amount of stores in the opposite direction, i.e. the algorithmic 16*L bytes for a dense
// 1-qubit gate on a global qubit.
//
// NCCL is used for the bootstrap (all-gather of IPC handles) and the scalar collectives.
#include <dlfcn.h>
#include <nccl.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>

#include "iqsb_internal.cuh"

// NCCL is resolved at run time, and only when a job has more than one rank: a single-rank process
// never loads it, and a process that already holds an NCCL (e.g. the one bundled with PyTorch) keeps
// using that copy instead of pulling a second libnccl.so.2 into the address space.
namespace nccl_dyn {
ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
const char *(*GetErrorString)(ncclResult_t) = nullptr;
static bool load() {
  if (GetUniqueId) return true;
  void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) {
    iqsb_set_error("cannot load libnccl.so.2: %s", dlerror());
    return false;
  }
#define IQSB_SYM(field, name)                                  \
  *(void **)(&field) = dlsym(h, name);                         \
  if (!field) {                                                \
    iqsb_set_error("libnccl.so.2 has no symbol %s", name);     \
    GetUniqueId = nullptr;                                     \
    return false;                                              \
  }
  IQSB_SYM(CommInitRank, "ncclCommInitRank")
  IQSB_SYM(CommDestroy, "ncclCommDestroy")
  IQSB_SYM(AllGather, "ncclAllGather")
  IQSB_SYM(AllReduce, "ncclAllReduce")
  IQSB_SYM(Broadcast, "ncclBroadcast")
  IQSB_SYM(GetErrorString, "ncclGetErrorString")
  IQSB_SYM(GetUniqueId, "ncclGetUniqueId")
#undef IQSB_SYM
  return true;
}
}  // namespace nccl_dyn
#define ncclGetUniqueId nccl_dyn::GetUniqueId
#define ncclCommInitRank nccl_dyn::CommInitRank
#define ncclCommDestroy nccl_dyn::CommDestroy
#define ncclAllGather nccl_dyn::AllGather
#define ncclAllReduce nccl_dyn::AllReduce
#define ncclBroadcast nccl_dyn::Broadcast
#define ncclGetErrorString nccl_dyn::GetErrorString

#define IQSB_NCCL(call)                                                                      \
  do {                                                                                       \
    ncclResult_t r__ = (call);                                                               \
    if (r__ != ncclSuccess) {                                                                \
      iqsb_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, ncclGetErrorString(r__)); \
      return IQSB_ERR_NCCL;                                                                  \
    }                                                                                        \
  } while (0)

struct iqsb_peer_table {
  ncclComm_t comm = nullptr;
  uint32_t *flags_local = nullptr;  // [nranks + 1] on this GPU; slot r is written by rank r, slot nranks is the abort word
  uint32_t **flags_peer = nullptr;  // host array: peer-mapped pointer to every rank's flags
  uint32_t **d_flags_peer = nullptr;
  uint32_t epoch = 0;
  double *d_coll = nullptr;  // staging for scalar collectives
  void *d_handles = nullptr; // [nranks * 64] staging for handle all-gathers
};

namespace {

constexpr int kMaxCollDoubles = 64;

__device__ __forceinline__ uint64_t global_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Every rank stores `epoch` into slot `me` of every peer's flag array, then waits until all
// of its own slots reached `epoch`.  Kernel boundaries on the stream order it after the
// preceding gate kernel; the system fence publishes that kernel's peer stores first.
//
// A rank that died (or never arrives) must not hang the other GPUs for ever: the wait has a
// wall-clock deadline.  The rank whose deadline passes raises the abort word (slot nranks) on every
// rank and records IQSB_ERR_PEER in its status word; every spinning rank sees its abort word, stops
// and records the same.  From then on barriers return at once -- the ordering guarantee is gone and
// the host gets the error from the next call that synchronises (iqsb_sync, iqsb_barrier, any
// reduction; iqsb_check at any time).
__global__ void k_barrier(uint32_t *const *peer_flags, volatile uint32_t *my_flags, int me, int nranks, uint32_t epoch, uint64_t timeout_ns,
                          volatile int *status) {
  int r = threadIdx.x;
  if (r < nranks && r != me) {
    __threadfence_system();
    volatile uint32_t *dst = peer_flags[r] + me;
    *dst = epoch;
    __threadfence_system();
    const uint64_t t0 = global_ns();
    unsigned spins = 0;
    bool failed = false;
    while ((int32_t)(my_flags[r] - epoch) < 0) {
      if (my_flags[nranks] != 0u) { failed = true; break; }  // somebody gave up
      if ((++spins & 1023u) == 0u && timeout_ns != 0 && global_ns() - t0 > timeout_ns) {
        for (int q = 0; q < nranks; ++q) *(volatile uint32_t *)(peer_flags[q] + nranks) = 1u + (uint32_t)r;  // who was missing, +1
        __threadfence_system();
        failed = true;
        break;
      }
    }
    if (failed) *status = IQSB_ERR_PEER;
    __threadfence_system();
  }
}

template <typename T>
__global__ void __launch_bounds__(256) k_copy_chunks(const Chunk<T> *__restrict__ src, Chunk<T> *__restrict__ dst, uint64_t n) {
  const uint64_t stride = (uint64_t)gridDim.x * 256;
  for (uint64_t t = (uint64_t)blockIdx.x * 256 + threadIdx.x; t < n; t += stride) st_chunk(dst + t, ld_chunk(src + t));
}

// exchange two equally sized chunk ranges (one may be remote): pure data movement
template <typename T>
__global__ void __launch_bounds__(256) k_xchg_chunks(Chunk<T> *a, Chunk<T> *b, uint64_t n) {
  const uint64_t stride = (uint64_t)gridDim.x * 256;
  uint64_t t = (uint64_t)blockIdx.x * 256 + threadIdx.x;
  for (; t + stride < n; t += 2 * stride) {
    Chunk<T> x0 = ld_chunk(a + t), y0 = ld_chunk(b + t);
    Chunk<T> x1 = ld_chunk(a + t + stride), y1 = ld_chunk(b + t + stride);
    st_chunk(a + t, y0);
    st_chunk(b + t, x0);
    st_chunk(a + t + stride, y1);
    st_chunk(b + t + stride, x1);
  }
  for (; t < n; t += stride) {
    Chunk<T> x = ld_chunk(a + t), y = ld_chunk(b + t);
    st_chunk(a + t, y);
    st_chunk(b + t, x);
  }
}

}  // namespace

static int allgather_bytes64(iqsb_ctx *ctx, const void *mine64, void *all_host) {
  iqsb_peer_table *pt = ctx->peers;
  IQSB_CUDA(cudaMemcpyAsync((char *)pt->d_handles + 64 * ctx->rank, mine64, 64, cudaMemcpyHostToDevice, ctx->stream));
  IQSB_NCCL(ncclAllGather((char *)pt->d_handles + 64 * ctx->rank, pt->d_handles, 64, ncclChar, pt->comm, ctx->stream));
  IQSB_CUDA(cudaMemcpyAsync(all_host, pt->d_handles, 64 * ctx->nranks, cudaMemcpyDeviceToHost, ctx->stream));
  IQSB_CUDA(cudaStreamSynchronize(ctx->stream));
  return IQSB_OK;
}

static int open_peers(iqsb_ctx *ctx, void *mine, void **out_ptrs) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t h;
  IQSB_CUDA(cudaIpcGetMemHandle(&h, mine));
  cudaIpcMemHandle_t *all = new cudaIpcMemHandle_t[ctx->nranks];
  int rc = allgather_bytes64(ctx, &h, all);
  if (rc != IQSB_OK) { delete[] all; return rc; }
  for (int r = 0; r < ctx->nranks; ++r) {
    if (r == ctx->rank) { out_ptrs[r] = mine; continue; }
    cudaError_t e = cudaIpcOpenMemHandle(&out_ptrs[r], all[r], cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      iqsb_set_error("cudaIpcOpenMemHandle(rank %d) failed: %s", r, cudaGetErrorString(e));
      delete[] all;
      return IQSB_ERR_CUDA;
    }
  }
  delete[] all;
  return IQSB_OK;
}

static void trace_init(const char *what) {  // IQS_B200_TRACE=1: bootstrap phase timestamps on stderr
  static const bool on = getenv("IQS_B200_TRACE") != nullptr;
  if (!on) return;
  struct timeval tv;
  gettimeofday(&tv, nullptr);
  fprintf(stderr, "[iqsb %ld.%03ld] %s\n", (long)tv.tv_sec % 1000, (long)tv.tv_usec / 1000, what);
}

// The ranks of one job live on one NVSwitch box and NCCL only carries the bootstrap and a few
// scalars, so unless the user says otherwise it is told to rendezvous over loopback and not to
// probe InfiniBand / network plugins (measured: the probing made ncclCommInitRank take 25-185 s on
// the pool's boxes; with these defaults it takes about a second).
static void nccl_single_node_defaults() {
  static bool done = false;
  if (done) return;
  done = true;
  setenv("NCCL_SOCKET_IFNAME", "lo", 0);
  setenv("NCCL_IB_DISABLE", "1", 0);
  setenv("NCCL_NET_PLUGIN", "none", 0);
  setenv("NCCL_TUNER_PLUGIN", "none", 0);
}

extern "C" int iqsb_unique_id(void *out_128_bytes) {
  IQSB_REQUIRE(out_128_bytes, "iqsb_unique_id: null argument");
  nccl_single_node_defaults();
  if (!nccl_dyn::load()) return IQSB_ERR_NCCL;
  static_assert(sizeof(ncclUniqueId) == IQSB_UNIQUE_ID_BYTES, "ncclUniqueId size");
  ncclUniqueId id;
  IQSB_NCCL(ncclGetUniqueId(&id));
  memcpy(out_128_bytes, &id, sizeof(id));
  return IQSB_OK;
}

int iqsb_comm_init(iqsb_ctx *ctx, const void *uid) {
  IQSB_REQUIRE(ctx->nranks <= 32, "iqsb_init: at most 32 ranks (the barrier kernel signals one peer per lane), %d requested", ctx->nranks);
  iqsb_peer_table *pt = new iqsb_peer_table();
  ctx->peers = pt;
  ncclUniqueId id;
  memcpy(&id, uid, sizeof(id));
  nccl_single_node_defaults();
  if (!nccl_dyn::load()) return IQSB_ERR_NCCL;
  trace_init("ncclCommInitRank ...");
  IQSB_NCCL(ncclCommInitRank(&pt->comm, ctx->nranks, id, ctx->rank));
  trace_init("ncclCommInitRank done");
  ctx->comm = pt->comm;
  IQSB_CUDA(cudaMalloc(&pt->d_coll, sizeof(double) * kMaxCollDoubles));
  IQSB_CUDA(cudaMalloc(&pt->d_handles, 64 * ctx->nranks));
  IQSB_CUDA(cudaMalloc(&pt->flags_local, sizeof(uint32_t) * (ctx->nranks + 1)));
  IQSB_CUDA(cudaMemset(pt->flags_local, 0, sizeof(uint32_t) * (ctx->nranks + 1)));
  if (const char *e = getenv("IQS_B200_BARRIER_TIMEOUT_S")) ctx->barrier_timeout_s = atof(e);
  pt->flags_peer = new uint32_t *[ctx->nranks];
  IQSB_TRY(open_peers(ctx, pt->flags_local, (void **)pt->flags_peer));
  IQSB_CUDA(cudaMalloc(&pt->d_flags_peer, sizeof(uint32_t *) * ctx->nranks));
  IQSB_CUDA(cudaMemcpy(pt->d_flags_peer, pt->flags_peer, sizeof(uint32_t *) * ctx->nranks, cudaMemcpyHostToDevice));
  trace_init("peer flags mapped");
  // nobody may signal before everyone has zeroed and mapped the flags
  IQSB_NCCL(ncclAllReduce(pt->d_coll, pt->d_coll, 1, ncclDouble, ncclSum, pt->comm, ctx->stream));
  IQSB_CUDA(cudaStreamSynchronize(ctx->stream));
  trace_init("first all-reduce done");
  return IQSB_OK;
}

int iqsb_comm_finalize(iqsb_ctx *ctx) {
  iqsb_peer_table *pt = ctx->peers;
  if (!pt) return IQSB_OK;
  for (int r = 0; r < ctx->nranks; ++r)
    if (r != ctx->rank && pt->flags_peer && pt->flags_peer[r]) cudaIpcCloseMemHandle(pt->flags_peer[r]);
  if (pt->comm) ncclCommDestroy(pt->comm);
  cudaFree(pt->d_coll);
  cudaFree(pt->d_handles);
  cudaFree(pt->flags_local);
  cudaFree(pt->d_flags_peer);
  delete[] pt->flags_peer;
  delete pt;
  ctx->peers = nullptr;
  return IQSB_OK;
}

int iqsb_peer_barrier(iqsb_ctx *ctx);  // also used by exchange.cu
static int peer_barrier(iqsb_ctx *ctx) { return iqsb_peer_barrier(ctx); }
int iqsb_peer_barrier(iqsb_ctx *ctx) {
  iqsb_peer_table *pt = ctx->peers;
  pt->epoch++;
  const uint64_t timeout_ns = ctx->barrier_timeout_s > 0. ? (uint64_t)(ctx->barrier_timeout_s * 1e9) : 0ull;
  k_barrier<<<1, 32, 0, ctx->stream>>>(pt->d_flags_peer, pt->flags_local, ctx->rank, ctx->nranks, pt->epoch, timeout_ns, ctx->d_status);
  return iqsb_check_launch(ctx, "k_barrier");
}

extern "C" int iqsb_barrier(iqsb_ctx *ctx) {
  IQSB_REQUIRE(ctx, "iqsb_barrier: null context");
  if (ctx->nranks > 1) IQSB_TRY(peer_barrier(ctx));
  IQSB_CUDA(cudaStreamSynchronize(ctx->stream));
  return iqsb_check(ctx);
}

extern "C" int iqsb_allreduce_f64(iqsb_ctx *ctx, double *inout, int n, int op) {
  IQSB_REQUIRE(ctx && inout, "iqsb_allreduce_f64: null argument");
  IQSB_REQUIRE(n >= 0 && n <= kMaxCollDoubles, "iqsb_allreduce_f64: at most %d values", kMaxCollDoubles);
  if (ctx->nranks == 1 || n == 0) return IQSB_OK;
  iqsb_peer_table *pt = ctx->peers;
  IQSB_CUDA(cudaMemcpyAsync(pt->d_coll, inout, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  IQSB_NCCL(ncclAllReduce(pt->d_coll, pt->d_coll, n, ncclDouble, op == IQSB_MAX ? ncclMax : ncclSum, pt->comm, ctx->stream));
  IQSB_CUDA(cudaMemcpyAsync(inout, pt->d_coll, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  IQSB_CUDA(cudaStreamSynchronize(ctx->stream));
  return IQSB_OK;
}

extern "C" int iqsb_bcast_f64(iqsb_ctx *ctx, double *inout, int n, int root) {
  IQSB_REQUIRE(ctx && inout, "iqsb_bcast_f64: null argument");
  IQSB_REQUIRE(n >= 0 && n <= kMaxCollDoubles, "iqsb_bcast_f64: at most %d values", kMaxCollDoubles);
  IQSB_REQUIRE(root >= 0 && root < ctx->nranks, "iqsb_bcast_f64: bad root");
  if (ctx->nranks == 1 || n == 0) return IQSB_OK;
  iqsb_peer_table *pt = ctx->peers;
  IQSB_CUDA(cudaMemcpyAsync(pt->d_coll, inout, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  IQSB_NCCL(ncclBroadcast(pt->d_coll, pt->d_coll, n, ncclDouble, root, pt->comm, ctx->stream));
  IQSB_CUDA(cudaMemcpyAsync(inout, pt->d_coll, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  IQSB_CUDA(cudaStreamSynchronize(ctx->stream));
  return IQSB_OK;
}

extern "C" int iqsb_share(iqsb_state *st) {
  IQSB_REQUIRE(st, "iqsb_share: null argument");
  iqsb_ctx *ctx = st->ctx;
  if (st->shared) return IQSB_OK;
  st->peer_ptr = new void *[ctx->nranks];
  for (int r = 0; r < ctx->nranks; ++r) st->peer_ptr[r] = nullptr;
  st->peer_ptr[ctx->rank] = st->d;
  if (ctx->nranks > 1) {
    IQSB_REQUIRE(st->mem_kind == IQSB_MEM_DEVICE, "iqsb_share: only IQSB_MEM_DEVICE shards can be shared");
    IQSB_TRY(open_peers(ctx, st->d, st->peer_ptr));
  }
  st->shared = true;
  return IQSB_OK;
}

int iqsb_comm_unshare(iqsb_state *st) {
  iqsb_ctx *ctx = st->ctx;
  if (ctx->nranks > 1) {
    // peers may still be reading this shard: rendezvous before unmapping / freeing
    peer_barrier(ctx);
    cudaStreamSynchronize(ctx->stream);
    for (int r = 0; r < ctx->nranks; ++r)
      if (r != ctx->rank && st->peer_ptr[r]) cudaIpcCloseMemHandle(st->peer_ptr[r]);
    peer_barrier(ctx);
    cudaStreamSynchronize(ctx->stream);
  }
  delete[] st->peer_ptr;
  st->peer_ptr = nullptr;
  st->shared = false;
  return IQSB_OK;
}

// ---------------------------------------------------------------------------------------
// global-qubit gates
// ---------------------------------------------------------------------------------------
// Who updates which pairs is decided by a pure host function, iqsb_plan_global, so that the
// partition logic can be exercised without a GPU (tests/test_partition_gloo.py).
namespace {

// highest local bit that is not in `avoid` (or -1)
int pick_split_bit(unsigned M, int avoid) {
  for (int b = (int)M - 1; b >= 0; --b)
    if (b != avoid) return b;
  return -1;
}

void add_fix(iqsb_plan *p, unsigned pos, unsigned val) {
  p->pos[p->nfix] = pos;
  p->val[p->nfix] = val;
  p->nfix++;
}

// Launch the two-pointer pair kernel for a plan; s0/s1 are shard base pointers (possibly remote).
int launch_plan(iqsb_state *st, const iqsb_plan &pl, void *s0, void *s1, const double m[8]) {
  unsigned p[3], v[3];
  int nfix = pl.nfix;
  for (int i = 0; i < nfix; ++i) { p[i] = pl.pos[i]; v[i] = pl.val[i]; }
  for (int i = 0; i < nfix; ++i)
    for (int j = i + 1; j < nfix; ++j)
      if (p[j] < p[i]) { unsigned t = p[i]; p[i] = p[j]; p[j] = t; t = v[i]; v[i] = v[j]; v[j] = t; }
  uint64_t L = st->local_amps;
  bool w1 = (nfix > 0 && p[0] == 0) || ((pl.extra0 | pl.extra1) & 1) || L < 2;
  unsigned ins[3];
  uint64_t fixed = 0;
  if (w1) {
    for (int i = 0; i < nfix; ++i) { ins[i] = p[i]; fixed |= (uint64_t)v[i] << p[i]; }
    Geom g = make_geom(L >> nfix, nfix, ins, fixed + pl.extra0, fixed + pl.extra1);
    return iqsb_launch_pairs(st, s0, s1, 1, g, m);
  }
  for (int i = 0; i < nfix; ++i) { ins[i] = p[i] - 1; fixed |= (uint64_t)v[i] << (p[i] - 1); }
  Geom g = make_geom((L / 2) >> nfix, nfix, ins, fixed + pl.extra0 / 2, fixed + pl.extra1 / 2);
  return iqsb_launch_pairs(st, s0, s1, 2, g, m);
}

int run_global(iqsb_state *st, int kind, unsigned M, unsigned pos1, unsigned pos2, const double m[8], const char *what) {
  iqsb_ctx *ctx = st->ctx;
  IQSB_REQUIRE(ctx->nranks > 1 && st->shared, "%s: register is not shared across ranks", what);
  IQSB_REQUIRE(M == st->log2_local, "%s: M must equal log2(local_amps)", what);
  iqsb_plan pl;
  IQSB_TRY(iqsb_plan_global(kind, ctx->rank, ctx->nranks, M, pos1, pos2, &pl));
  IQSB_TRY(peer_barrier(ctx));  // the partner's earlier kernels are complete
  if (pl.active) {
    void *mine = st->d, *theirs = st->peer_ptr[pl.partner];
    IQSB_TRY(launch_plan(st, pl, pl.role == 0 ? mine : theirs, pl.role == 0 ? theirs : mine, m));
  }
  ctx->nvlink_bytes += pl.link_amps * st->amp_bytes();
  return peer_barrier(ctx);  // the partner's stores into this shard are complete
}

}  // namespace

// kind 0: 1-qubit gate on global position pos2 (pos1 ignored)        -- HP_Distrpair(P)
// kind 1: controlled gate, control pos1 local, target pos2 global    -- HP_Distrpair(C,T)
// kind 2: swap-family gate, pos1 < pos2, pos2 global                 -- HP_DistrSwap
extern "C" int iqsb_plan_global(int kind, int rank, int nranks, unsigned M, unsigned pos1, unsigned pos2, iqsb_plan *out) {
  IQSB_REQUIRE(out, "iqsb_plan_global: null plan");
  IQSB_REQUIRE(nranks > 1 && (nranks & (nranks - 1)) == 0 && rank >= 0 && rank < nranks, "iqsb_plan_global: bad rank/nranks");
  IQSB_REQUIRE(kind >= 0 && kind <= 2, "iqsb_plan_global: bad kind");
  IQSB_REQUIRE(pos2 >= M && (1u << (pos2 - M)) < (unsigned)nranks, "iqsb_plan_global: position %u is not a global position", pos2);
  memset(out, 0, sizeof(*out));
  const uint64_t L = 1ull << M;
  if (kind == 0 || kind == 1 || (kind == 2 && pos1 < M)) {
    if (kind == 1) IQSB_REQUIRE(pos1 < M, "iqsb_plan_global: the control must be local");
    if (kind == 2) IQSB_REQUIRE(pos1 < pos2, "iqsb_plan_global: need pos1 < pos2");
    // the pair (a on the rank whose pos2 bit is 0, b on its partner) is owned by exactly one of the two
    // ranks, chosen by one local split bit: bit value 1 -> the 0-side rank (the reference's i-task takes
    // the upper half, 1q.cpp:130-152), 0 -> the 1-side rank.
    unsigned rb = pos2 - M;
    unsigned mybit = (rank >> rb) & 1;
    out->partner = rank ^ (1 << rb);
    out->role = (int)mybit;
    int avoid = kind == 0 ? -1 : (int)pos1;
    if (kind == 1) add_fix(out, pos1, 1);  // control bit set
    if (kind == 2) {                       // i0 = base + 2^pos1 on the 0-side, i1 = base on the 1-side
      add_fix(out, pos1, 0);
      out->extra0 = 1ull << pos1;
    }
    int sb = pick_split_bit(M, avoid);
    if (sb < 0) out->active = mybit == 0;  // a single pair: the 0-side rank does it
    else {
      add_fix(out, (unsigned)sb, mybit ? 0u : 1u);
      out->active = 1;
    }
    out->npairs = out->active ? (L >> out->nfix) : 0;
    out->link_amps = kind == 0 ? L : L / 2;  // per direction, SURVEY.md 8d
  } else {
    // both positions global: ranks with (bit1, bit2) = (1,0) exchange with (0,1); equal bits idle
    IQSB_REQUIRE(pos1 < pos2, "iqsb_plan_global: need pos1 < pos2");
    unsigned r1 = pos1 - M, r2 = pos2 - M;
    unsigned b1 = (rank >> r1) & 1, b2 = (rank >> r2) & 1;
    if (b1 != b2) {
      out->partner = rank ^ (1 << r1) ^ (1 << r2);
      out->role = b1 == 1 ? 0 : 1;  // the rank holding i0 (pos1 bit 1, pos2 bit 0) is the 0-side
      int sb = pick_split_bit(M, -1);
      if (sb < 0) out->active = out->role == 0;
      else {
        add_fix(out, (unsigned)sb, out->role == 0 ? 1u : 0u);
        out->active = 1;
      }
      out->npairs = out->active ? (L >> out->nfix) : 0;
      out->link_amps = L;
    }
  }
  return IQSB_OK;
}

extern "C" int iqsb_gate1_global(iqsb_state *st, unsigned M, unsigned pos, const double m[8]) {
  IQSB_REQUIRE(st && m, "iqsb_gate1_global: null argument");
  return run_global(st, 0, M, 0, pos, m, "iqsb_gate1_global");
}

extern "C" int iqsb_cgate1_global(iqsb_state *st, unsigned M, unsigned cpos, unsigned tpos, const double m[8]) {
  IQSB_REQUIRE(st && m, "iqsb_cgate1_global: null argument");
  return run_global(st, 1, M, cpos, tpos, m, "iqsb_cgate1_global");
}

extern "C" int iqsb_swap2x2_global(iqsb_state *st, unsigned M, unsigned pos1, unsigned pos2, const double m[8]) {
  IQSB_REQUIRE(st && m, "iqsb_swap2x2_global: null argument");
  return run_global(st, 2, M, pos1, pos2, m, "iqsb_swap2x2_global");
}

extern "C" int iqsb_idle_global(iqsb_state *st) {
  IQSB_REQUIRE(st, "iqsb_idle_global: null argument");
  iqsb_ctx *ctx = st->ctx;
  IQSB_REQUIRE(ctx->nranks > 1 && st->shared, "iqsb_idle_global: register is not shared across ranks");
  IQSB_TRY(peer_barrier(ctx));
  return peer_barrier(ctx);
}

// pairwise < 0: not known -- the ranks agree through a 1-double all-reduce (iqsb_permute_global, where
// each rank only knows its own source and destination); 0 / 1: decided by the caller from the
// permutation of the rank bits, which every rank knows (iqsb_permute_global_bits): no communication.
static int permute_global_impl(iqsb_state *st, int src_rank, int dst_rank, int pairwise) {
  iqsb_ctx *ctx = st->ctx;
  IQSB_REQUIRE(ctx->nranks > 1 && st->shared, "iqsb_permute_global: register is not shared across ranks");
  IQSB_REQUIRE(src_rank >= 0 && src_rank < ctx->nranks && dst_rank >= 0 && dst_rank < ctx->nranks, "iqsb_permute_global: bad ranks");
  uint64_t Lall = st->local_amps;
  // Fast path: when every rank is a fixed point or half of a 2-cycle (e.g. reversing the order of the
  // global qubits), partners exchange their shards in place with one kernel each -- the lower rank
  // swaps the upper half, the higher rank the lower half -- without staging.  All ranks must take the
  // same path (the rendezvous counts differ).
  {
    double not_simple = (src_rank == dst_rank) ? 0.0 : 1.0;
    if (pairwise < 0) IQSB_TRY(iqsb_allreduce_f64(ctx, &not_simple, 1, IQSB_MAX));
    else not_simple = pairwise ? 0.0 : 1.0;
    if (not_simple == 0.0 && Lall >= 4) {
      IQSB_TRY(peer_barrier(ctx));
      if (src_rank != ctx->rank) {
        uint64_t half_chunks = Lall / 4;  // chunks in half a shard
        uint64_t first = ctx->rank < src_rank ? half_chunks : 0;
        char *mine = (char *)st->d + first * 2 * st->amp_bytes();
        char *theirs = (char *)st->peer_ptr[src_rank] + first * 2 * st->amp_bytes();
        int grid = ctx->num_sms * 8;
        if (st->dtype == IQSB_F64) k_xchg_chunks<double><<<grid, 256, 0, ctx->stream>>>((Chunk<double> *)mine, (Chunk<double> *)theirs, half_chunks);
        else k_xchg_chunks<float><<<grid, 256, 0, ctx->stream>>>((Chunk<float> *)mine, (Chunk<float> *)theirs, half_chunks);
        IQSB_TRY(iqsb_check_launch(ctx, "k_xchg_chunks"));
        ctx->nvlink_bytes += Lall * st->amp_bytes();
      }
      return peer_barrier(ctx);
    }
  }
  // Every shard is both a source and a destination, so the move is staged through the tmp area
  // in chunks (as the reference does, qureg_permute.cpp:174-185), pulling with peer loads.
  uint64_t L = st->local_amps, chunk = st->tmp_amps;
  void *stage = (char *)st->d + L * st->amp_bytes();
  void *owned = nullptr;
  if (chunk == 0) {
    chunk = L < (1ull << 26) ? L : (1ull << 26);
    IQSB_CUDA(cudaMalloc(&owned, chunk * st->amp_bytes()));
    stage = owned;
  }
  if (chunk > L) chunk = L;
  IQSB_REQUIRE(chunk >= 2 && L % chunk == 0, "iqsb_permute_global: tmp size must divide the shard");
  int rc = IQSB_OK;
  int grid = ctx->num_sms * 8;
  for (uint64_t c = 0; c < L && rc == IQSB_OK; c += chunk) {
    rc = peer_barrier(ctx);  // sources are final up to here
    if (rc != IQSB_OK) break;
    const char *src = (const char *)st->peer_ptr[src_rank] + c * st->amp_bytes();
    char *mine = (char *)st->d + c * st->amp_bytes();
    if (st->dtype == IQSB_F64)
      k_copy_chunks<double><<<grid, 256, 0, ctx->stream>>>((const Chunk<double> *)src, (Chunk<double> *)stage, chunk / 2);
    else
      k_copy_chunks<float><<<grid, 256, 0, ctx->stream>>>((const Chunk<float> *)src, (Chunk<float> *)stage, chunk / 2);
    rc = iqsb_check_launch(ctx, "k_copy_chunks");
    if (rc != IQSB_OK) break;
    rc = peer_barrier(ctx);  // everybody has pulled chunk c: it may be overwritten
    if (rc != IQSB_OK) break;
    if (st->dtype == IQSB_F64)
      k_copy_chunks<double><<<grid, 256, 0, ctx->stream>>>((const Chunk<double> *)stage, (Chunk<double> *)mine, chunk / 2);
    else
      k_copy_chunks<float><<<grid, 256, 0, ctx->stream>>>((const Chunk<float> *)stage, (Chunk<float> *)mine, chunk / 2);
    rc = iqsb_check_launch(ctx, "k_copy_chunks");
  }
  if (rc == IQSB_OK && src_rank != ctx->rank) ctx->nvlink_bytes += L * st->amp_bytes();
  if (rc == IQSB_OK) rc = peer_barrier(ctx);
  if (owned) {
    cudaStreamSynchronize(ctx->stream);
    cudaFree(owned);
  }
  return rc;
}

extern "C" int iqsb_permute_global(iqsb_state *st, int src_rank, int dst_rank) {
  IQSB_REQUIRE(st, "iqsb_permute_global: null argument");
  return permute_global_impl(st, src_rank, dst_rank, -1);
}

// Pure host function (no GPU needed): what `rank` does when the content of rank bit b moves to rank bit
// dst_rank_bit[b].  Every rank derives its own source and destination AND whether all ranks move in
// pairs (the bit permutation is an involution) from the same table: nothing has to be agreed at run time.
extern "C" int iqsb_plan_permute_global_bits(int rank, int nranks, const uint8_t *dst_rank_bit, unsigned nbits, int *source, int *destination,
                                             int *pairwise, int *identity) {
  IQSB_REQUIRE(dst_rank_bit && source && destination && pairwise && identity, "iqsb_plan_permute_global_bits: null argument");
  IQSB_REQUIRE(nbits < 31 && (1 << nbits) == nranks && rank >= 0 && rank < nranks, "iqsb_plan_permute_global_bits: %u rank bits do not describe rank %d of %d",
               nbits, rank, nranks);
  unsigned inverse[32], seen = 0;
  for (unsigned b = 0; b < nbits; ++b) {
    IQSB_REQUIRE(dst_rank_bit[b] < nbits && !((seen >> dst_rank_bit[b]) & 1u), "iqsb_plan_permute_global_bits: not a permutation of the rank bits");
    seen |= 1u << dst_rank_bit[b];
    inverse[dst_rank_bit[b]] = b;
  }
  bool involution = true, same = true;
  int dst = 0, src = 0;
  for (unsigned b = 0; b < nbits; ++b) {
    involution = involution && dst_rank_bit[dst_rank_bit[b]] == b;
    same = same && dst_rank_bit[b] == b;
    if ((rank >> b) & 1) dst |= 1 << dst_rank_bit[b];        // my bit b travels to bit dst[b]
    if ((rank >> dst_rank_bit[b]) & 1) src |= 1 << b;        // my bit dst[b] is filled from bit b of the source rank
  }
  (void)inverse;
  *source = src;
  *destination = dst;
  *pairwise = involution ? 1 : 0;
  *identity = same ? 1 : 0;
  return IQSB_OK;
}

extern "C" int iqsb_permute_global_bits(iqsb_state *st, const uint8_t *dst_rank_bit, unsigned nbits) {
  IQSB_REQUIRE(st && dst_rank_bit, "iqsb_permute_global_bits: null argument");
  iqsb_ctx *ctx = st->ctx;
  int source = 0, destination = 0, pairwise = 0, identity = 0;
  IQSB_TRY(iqsb_plan_permute_global_bits(ctx->rank, ctx->nranks, dst_rank_bit, nbits, &source, &destination, &pairwise, &identity));
  if (identity) return IQSB_OK;
  return permute_global_impl(st, source, destination, pairwise);
}
