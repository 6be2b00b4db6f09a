// api.cu -- context and memory management behind the C ABI (include/iqsb.h).
//
// Replaces: iqs::mpi::Environment bootstrap (reference src/mpi_env.cpp:95-230) and
// QubitRegister::Allocate/Resize/~QubitRegister (src/qureg_init.cpp:44-75, 148-187, 449-457):
// the state vector lives in HBM (cudaMalloc, or cudaMallocManaged with the device as the
// preferred location when the caller needs a host-dereferenceable pointer).
#include <stdarg.h>
#include <string.h>

#include <string>
#include <vector>

#include "iqsb_internal.cuh"

static thread_local char g_err[1024] = "";

void iqsb_set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// Per-kernel-class timing: while profiling is on, every launch is followed by a CUDA event on the
// engine's stream; the time between two consecutive events is the duration of the kernel between them
// (the stream is kept busy by the caller, so there is no idle time to mis-attribute; what little
// there is counts against the kernel, never for it).
struct iqsb_prof {
  bool on = false, overflow = false;
  std::vector<cudaEvent_t> pool;  // pool[0] = start marker, pool[i + 1] closes recs[i]
  struct Rec {
    const char *name;
    double bytes;
  };
  std::vector<Rec> recs;
};
static const size_t kMaxProfLaunches = 1u << 16;

int iqsb_check_launch(iqsb_ctx *ctx, const char *what, double algo_bytes) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    iqsb_set_error("launch of %s failed: %s", what, cudaGetErrorString(e));
    return IQSB_ERR_CUDA;
  }
  ctx->launches++;
  iqsb_prof *p = ctx->prof;
  if (p && p->on) {
    if (p->recs.size() >= kMaxProfLaunches) {
      p->overflow = true;
    } else {
      const size_t slot = p->recs.size() + 1;
      if (p->pool.size() <= slot) {
        cudaEvent_t ev;
        IQSB_CUDA(cudaEventCreate(&ev));
        p->pool.push_back(ev);
      }
      IQSB_CUDA(cudaEventRecord(p->pool[slot], ctx->stream));
      p->recs.push_back({what, algo_bytes});
    }
  }
  return IQSB_OK;
}

extern "C" int iqsb_profile(iqsb_ctx *ctx, int on) {
  IQSB_REQUIRE(ctx, "iqsb_profile: null context");
  if (!ctx->prof) ctx->prof = new iqsb_prof();
  iqsb_prof *p = ctx->prof;
  if (on) {
    p->recs.clear();
    p->overflow = false;
    if (p->pool.empty()) {
      cudaEvent_t ev;
      IQSB_CUDA(cudaEventCreate(&ev));
      p->pool.push_back(ev);
    }
    IQSB_CUDA(cudaEventRecord(p->pool[0], ctx->stream));
  }
  p->on = on != 0;
  return IQSB_OK;
}

// JSON text: {"overflow": false, "classes": [{"name": "...", "launches": n, "ms": t, "bytes": b}, ...]}
extern "C" int iqsb_profile_read(iqsb_ctx *ctx, char *out, size_t cap) {
  IQSB_REQUIRE(ctx && out && cap > 0, "iqsb_profile_read: null argument");
  out[0] = 0;
  iqsb_prof *p = ctx->prof;
  std::string js = "{\"overflow\": ";
  js += (p && p->overflow) ? "true" : "false";
  js += ", \"classes\": [";
  if (p && !p->recs.empty()) {
    IQSB_CUDA(cudaEventSynchronize(p->pool[p->recs.size()]));
    struct Agg {
      const char *name;
      uint64_t n;
      double ms, bytes;
    };
    std::vector<Agg> agg;
    for (size_t i = 0; i < p->recs.size(); ++i) {
      float ms = 0.f;
      IQSB_CUDA(cudaEventElapsedTime(&ms, p->pool[i], p->pool[i + 1]));
      size_t k = 0;
      for (; k < agg.size(); ++k)
        if (agg[k].name == p->recs[i].name || strcmp(agg[k].name, p->recs[i].name) == 0) break;
      if (k == agg.size()) agg.push_back({p->recs[i].name, 0, 0., 0.});
      agg[k].n++;
      agg[k].ms += ms;
      agg[k].bytes += p->recs[i].bytes;
    }
    char buf[256];
    for (size_t k = 0; k < agg.size(); ++k) {
      snprintf(buf, sizeof(buf), "%s{\"name\": \"%s\", \"launches\": %llu, \"ms\": %.6f, \"bytes\": %.1f}", k ? ", " : "", agg[k].name,
               (unsigned long long)agg[k].n, agg[k].ms, agg[k].bytes);
      js += buf;
    }
  }
  js += "]}";
  IQSB_REQUIRE(js.size() + 1 <= cap, "iqsb_profile_read: buffer of %zu bytes is too small (%zu needed)", cap, js.size() + 1);
  memcpy(out, js.c_str(), js.size() + 1);
  return IQSB_OK;
}

// comm.cu
int iqsb_comm_init(iqsb_ctx *ctx, const void *uid);
int iqsb_comm_finalize(iqsb_ctx *ctx);
int iqsb_comm_unshare(iqsb_state *st);

extern "C" int iqsb_version(void) { return 100; }
extern "C" const char *iqsb_last_error(void) { return g_err; }

extern "C" int iqsb_init(int rank, int nranks, const void *uid, int device, iqsb_ctx **out) {
  IQSB_REQUIRE(out, "iqsb_init: null out pointer");
  IQSB_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "iqsb_init: bad rank %d / %d", rank, nranks);
  IQSB_REQUIRE((nranks & (nranks - 1)) == 0, "iqsb_init: number of ranks must be a power of two");
  IQSB_REQUIRE(nranks == 1 || uid, "iqsb_init: nranks > 1 needs a unique id");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    // No CPU fallback: the engine is CUDA only.
    iqsb_set_error("iqsb_init: no CUDA device available (%s)", e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    (void)cudaGetLastError();
    return IQSB_ERR_CUDA;
  }
  if (device < 0) device = rank % ndev;
  IQSB_REQUIRE(device < ndev, "iqsb_init: device %d out of range (%d visible)", device, ndev);
  IQSB_CUDA(cudaSetDevice(device));
  iqsb_ctx *ctx = new iqsb_ctx();
  ctx->rank = rank;
  ctx->nranks = nranks;
  ctx->device = device;
  cudaDeviceProp prop;
  IQSB_CUDA(cudaGetDeviceProperties(&prop, device));
  ctx->num_sms = prop.multiProcessorCount;
  IQSB_CUDA(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
  ctx->stream = ctx->own_stream;
  IQSB_CUDA(cudaEventCreate(&ctx->ev0));
  IQSB_CUDA(cudaEventCreate(&ctx->ev1));
  IQSB_CUDA(cudaMalloc(&ctx->d_partials, sizeof(double) * kMaxRedBlocks * kMaxRedOut));
  IQSB_CUDA(cudaMalloc(&ctx->d_result, sizeof(double) * kMaxRedOut));
  IQSB_CUDA(cudaMalloc(&ctx->d_flags, sizeof(int) * 4));
  IQSB_CUDA(cudaMallocHost(&ctx->h_result, sizeof(double) * kMaxRedOut));
  IQSB_CUDA(cudaHostAlloc((void **)&ctx->h_status, sizeof(int), cudaHostAllocMapped));
  *ctx->h_status = IQSB_OK;
  IQSB_CUDA(cudaHostGetDevicePointer((void **)&ctx->d_status, ctx->h_status, 0));
  if (const char *a = getenv("IQS_B200_ARITH")) ctx->arith = (strcmp(a, "fma") == 0 || strcmp(a, "FMA") == 0) ? IQSB_ARITH_FMA : IQSB_ARITH_EXACT;
  if (nranks > 1) {
    int rc = iqsb_comm_init(ctx, uid);
    if (rc != IQSB_OK) {  // keep the message of the failure, release everything acquired so far
      std::string why = g_err;
      ctx->nranks = 1;  // nothing of the communication layer to tear down
      iqsb_finalize(ctx);
      iqsb_set_error("%s", why.c_str());
      return rc;
    }
  }
  *out = ctx;
  return IQSB_OK;
}

extern "C" int iqsb_finalize(iqsb_ctx *ctx) {
  if (!ctx) return IQSB_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->nranks > 1) iqsb_comm_finalize(ctx);
  cudaFree(ctx->d_partials);
  cudaFree(ctx->d_result);
  cudaFree(ctx->d_flags);
  cudaFreeHost(ctx->h_result);
  cudaFreeHost(ctx->h_status);
  cudaFree(ctx->d_tile_counter);
  if (ctx->slots) {
    for (int i = 0; i < kMaxEventSlots; ++i)
      if (ctx->slots[i]) cudaEventDestroy(ctx->slots[i]);
    delete[] ctx->slots;
  }
  if (ctx->prof) {
    for (cudaEvent_t ev : ctx->prof->pool) cudaEventDestroy(ev);
    delete ctx->prof;
  }
  cudaEventDestroy(ctx->ev0);
  cudaEventDestroy(ctx->ev1);
  cudaStreamDestroy(ctx->own_stream);
  delete ctx;
  return IQSB_OK;
}

extern "C" int iqsb_rank(const iqsb_ctx *ctx) { return ctx ? ctx->rank : 0; }
extern "C" int iqsb_nranks(const iqsb_ctx *ctx) { return ctx ? ctx->nranks : 1; }
extern "C" int iqsb_device(const iqsb_ctx *ctx) { return ctx ? ctx->device : -1; }
extern "C" uint64_t iqsb_launch_count(const iqsb_ctx *ctx) { return ctx ? ctx->launches : 0; }
extern "C" uint64_t iqsb_nvlink_bytes(const iqsb_ctx *ctx) { return ctx ? ctx->nvlink_bytes : 0; }

// What the kernels reported since the last call: today only the peer barrier, which gives up after
// IQS_B200_BARRIER_TIMEOUT_S seconds without a partner instead of spinning for ever.
extern "C" int iqsb_check(iqsb_ctx *ctx) {
  IQSB_REQUIRE(ctx, "iqsb_check: null context");
  const int st = ctx->h_status ? *(volatile int *)ctx->h_status : IQSB_OK;
  if (st != IQSB_OK) {
    iqsb_set_error("rank %d: a peer did not reach the barrier within %.0f s (or another rank gave up first); the ranks are out of step and the state is undefined",
                   ctx->rank, ctx->barrier_timeout_s);
    return st;
  }
  return IQSB_OK;
}

extern "C" int iqsb_sync(iqsb_ctx *ctx) {
  IQSB_REQUIRE(ctx, "iqsb_sync: null context");
  IQSB_CUDA(cudaStreamSynchronize(ctx->stream));
  return iqsb_check(ctx);
}

extern "C" int iqsb_mem_info(iqsb_ctx *ctx, uint64_t *free_bytes, uint64_t *total_bytes) {
  IQSB_REQUIRE(ctx && free_bytes && total_bytes, "iqsb_mem_info: null argument");
  size_t f = 0, t = 0;
  IQSB_CUDA(cudaSetDevice(ctx->device));
  IQSB_CUDA(cudaMemGetInfo(&f, &t));
  *free_bytes = f;
  *total_bytes = t;
  return IQSB_OK;
}

extern "C" int iqsb_set_stream(iqsb_ctx *ctx, void *cuda_stream) {
  IQSB_REQUIRE(ctx, "iqsb_set_stream: null context");
  IQSB_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
  return IQSB_OK;
}
extern "C" void *iqsb_get_stream(iqsb_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

extern "C" int iqsb_set_arith(iqsb_ctx *ctx, int mode) {
  IQSB_REQUIRE(ctx, "iqsb_set_arith: null context");
  IQSB_REQUIRE(mode == IQSB_ARITH_EXACT || mode == IQSB_ARITH_FMA, "iqsb_set_arith: unknown mode %d", mode);
  ctx->arith = mode;
  return IQSB_OK;
}
extern "C" int iqsb_get_arith(const iqsb_ctx *ctx) { return ctx ? ctx->arith : IQSB_ARITH_EXACT; }

extern "C" int iqsb_timer_start(iqsb_ctx *ctx) {
  IQSB_REQUIRE(ctx, "iqsb_timer_start: null context");
  IQSB_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  return IQSB_OK;
}
extern "C" int iqsb_timer_stop(iqsb_ctx *ctx, double *elapsed_ms) {
  IQSB_REQUIRE(ctx && elapsed_ms, "iqsb_timer_stop: null argument");
  IQSB_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  IQSB_CUDA(cudaEventSynchronize(ctx->ev1));
  float ms = 0.f;
  IQSB_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
  *elapsed_ms = (double)ms;
  return IQSB_OK;
}

extern "C" int iqsb_event_record(iqsb_ctx *ctx, int slot) {
  IQSB_REQUIRE(ctx && slot >= 0 && slot < kMaxEventSlots, "iqsb_event_record: bad slot");
  if (!ctx->slots) {
    ctx->slots = new cudaEvent_t[kMaxEventSlots];
    for (int i = 0; i < kMaxEventSlots; ++i) ctx->slots[i] = nullptr;
  }
  if (!ctx->slots[slot]) IQSB_CUDA(cudaEventCreate(&ctx->slots[slot]));
  IQSB_CUDA(cudaEventRecord(ctx->slots[slot], ctx->stream));
  return IQSB_OK;
}
extern "C" int iqsb_event_elapsed(iqsb_ctx *ctx, int slot_from, int slot_to, double *elapsed_ms) {
  IQSB_REQUIRE(ctx && elapsed_ms && ctx->slots && slot_from >= 0 && slot_from < kMaxEventSlots && slot_to >= 0 && slot_to < kMaxEventSlots &&
                   ctx->slots[slot_from] && ctx->slots[slot_to],
               "iqsb_event_elapsed: slots were not recorded");
  IQSB_CUDA(cudaEventSynchronize(ctx->slots[slot_to]));
  float ms = 0.f;
  IQSB_CUDA(cudaEventElapsedTime(&ms, ctx->slots[slot_from], ctx->slots[slot_to]));
  *elapsed_ms = (double)ms;
  return IQSB_OK;
}

// ---------------------------------------------------------------------------------------
// memory
// ---------------------------------------------------------------------------------------
extern "C" int iqsb_alloc(iqsb_ctx *ctx, uint64_t local_amps, uint64_t tmp_amps, int dtype, int mem_kind,
                          iqsb_state **out) {
  IQSB_REQUIRE(ctx && out, "iqsb_alloc: null argument");
  IQSB_REQUIRE(local_amps >= 1 && (local_amps & (local_amps - 1)) == 0, "iqsb_alloc: local_amps must be a power of two");
  IQSB_REQUIRE(dtype == IQSB_F64 || dtype == IQSB_F32, "iqsb_alloc: bad dtype");
  IQSB_REQUIRE(mem_kind == IQSB_MEM_DEVICE || mem_kind == IQSB_MEM_MANAGED, "iqsb_alloc: bad mem_kind");
  IQSB_CUDA(cudaSetDevice(ctx->device));
  iqsb_state *st = new iqsb_state();
  st->ctx = ctx;
  st->local_amps = local_amps;
  st->tmp_amps = tmp_amps;
  st->dtype = dtype;
  st->mem_kind = mem_kind;
  unsigned lg = 0;
  while ((1ull << lg) < local_amps) ++lg;
  st->log2_local = lg;
  size_t bytes = (size_t)(local_amps + tmp_amps) * st->amp_bytes();
  if (bytes < 32) bytes = 32;
  cudaError_t e;
  if (mem_kind == IQSB_MEM_MANAGED) {
    e = cudaMallocManaged(&st->d, bytes, cudaMemAttachGlobal);
    if (e == cudaSuccess) {
      // keep the pages in HBM; host accesses migrate on demand and iqsb_prefetch_device brings them back
      cudaMemAdvise(st->d, bytes, cudaMemAdviseSetPreferredLocation, ctx->device);
      cudaMemPrefetchAsync(st->d, bytes, ctx->device, ctx->stream);
      (void)cudaGetLastError();
    }
  } else {
    e = cudaMalloc(&st->d, bytes);
  }
  if (e != cudaSuccess) {
    iqsb_set_error("iqsb_alloc: cannot allocate %zu bytes: %s", bytes, cudaGetErrorString(e));
    (void)cudaGetLastError();
    delete st;
    return IQSB_ERR_CUDA;
  }
  *out = st;
  return IQSB_OK;
}

extern "C" int iqsb_free(iqsb_state *st) {
  if (!st) return IQSB_OK;
  cudaSetDevice(st->ctx->device);
  cudaStreamSynchronize(st->ctx->stream);
  if (st->shared) iqsb_comm_unshare(st);
  cudaFree(st->d);
  delete st;
  return IQSB_OK;
}

extern "C" uint64_t iqsb_local_amps(const iqsb_state *st) { return st ? st->local_amps : 0; }
extern "C" int iqsb_dtype(const iqsb_state *st) { return st ? st->dtype : -1; }
extern "C" void *iqsb_device_ptr(iqsb_state *st) { return st ? st->d : nullptr; }
extern "C" void *iqsb_host_ptr(iqsb_state *st) {
  if (!st || st->mem_kind != IQSB_MEM_MANAGED) return nullptr;
  cudaStreamSynchronize(st->ctx->stream);
  return st->d;
}

extern "C" int iqsb_prefetch_device(iqsb_state *st) {
  IQSB_REQUIRE(st, "iqsb_prefetch_device: null argument");
  if (st->mem_kind != IQSB_MEM_MANAGED) return IQSB_OK;
  size_t bytes = (size_t)(st->local_amps + st->tmp_amps) * st->amp_bytes();
  IQSB_CUDA(cudaMemPrefetchAsync(st->d, bytes, st->ctx->device, st->ctx->stream));
  return IQSB_OK;
}

extern "C" int iqsb_upload(iqsb_state *st, const void *host_amps, uint64_t first_amp, uint64_t count) {
  IQSB_REQUIRE(st && host_amps, "iqsb_upload: null argument");
  IQSB_REQUIRE(first_amp + count <= st->local_amps + st->tmp_amps, "iqsb_upload: range out of bounds");
  IQSB_CUDA(cudaMemcpyAsync((char *)st->d + first_amp * st->amp_bytes(), host_amps, count * st->amp_bytes(),
                            cudaMemcpyHostToDevice, st->ctx->stream));
  IQSB_CUDA(cudaStreamSynchronize(st->ctx->stream));
  return IQSB_OK;
}

extern "C" int iqsb_download(iqsb_state *st, void *host_amps, uint64_t first_amp, uint64_t count) {
  IQSB_REQUIRE(st && host_amps, "iqsb_download: null argument");
  IQSB_REQUIRE(first_amp + count <= st->local_amps + st->tmp_amps, "iqsb_download: range out of bounds");
  IQSB_CUDA(cudaMemcpyAsync(host_amps, (char *)st->d + first_amp * st->amp_bytes(), count * st->amp_bytes(),
                            cudaMemcpyDeviceToHost, st->ctx->stream));
  IQSB_CUDA(cudaStreamSynchronize(st->ctx->stream));
  return IQSB_OK;
}

extern "C" int iqsb_copy(iqsb_state *dst, const iqsb_state *src) {
  IQSB_REQUIRE(dst && src, "iqsb_copy: null argument");
  IQSB_REQUIRE(dst->local_amps == src->local_amps && dst->dtype == src->dtype, "iqsb_copy: registers do not match");
  IQSB_CUDA(cudaMemcpyAsync(dst->d, src->d, src->local_amps * src->amp_bytes(), cudaMemcpyDeviceToDevice,
                            dst->ctx->stream));
  return IQSB_OK;
}

extern "C" int iqsb_set_amp(iqsb_state *st, uint64_t local_index, double re, double im) {
  IQSB_REQUIRE(st, "iqsb_set_amp: null argument");
  IQSB_REQUIRE(local_index < st->local_amps, "iqsb_set_amp: index out of range");
  if (st->dtype == IQSB_F64) {
    double v[2] = {re, im};
    return iqsb_upload(st, v, local_index, 1);
  }
  float v[2] = {(float)re, (float)im};
  return iqsb_upload(st, v, local_index, 1);
}

extern "C" int iqsb_get_amp(iqsb_state *st, uint64_t local_index, double *re, double *im) {
  IQSB_REQUIRE(st && re && im, "iqsb_get_amp: null argument");
  IQSB_REQUIRE(local_index < st->local_amps, "iqsb_get_amp: index out of range");
  if (st->dtype == IQSB_F64) {
    double v[2];
    IQSB_TRY(iqsb_download(st, v, local_index, 1));
    *re = v[0];
    *im = v[1];
  } else {
    float v[2];
    IQSB_TRY(iqsb_download(st, v, local_index, 1));
    *re = v[0];
    *im = v[1];
  }
  return IQSB_OK;
}
