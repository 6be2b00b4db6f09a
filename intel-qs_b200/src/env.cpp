// env.cpp -- iqs::mpi::Environment over the C ABI context (replaces reference src/mpi_env.cpp, the
// MPI bootstrap, by an NCCL bootstrap: see include/mpi_env.hpp), plus the small free functions of
// utils.cpp / gate_spec.cpp.
#include <fcntl.h>
#include <sys/time.h>

#include <cassert>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <cstring>
#include <iostream>
#include <string>
#include <thread>
#include <vector>

#include "../include/bitops.hpp"
#include "../include/conversion.hpp"
#include "../include/gate_spec.hpp"
#include "../include/mpi_env.hpp"
#include "../include/mpi_utils.hpp"
#include "../include/utils.hpp"
#include "iqsb.h"

namespace iqs {

double time_in_seconds(void) {
  struct timeval tv;
  gettimeofday(&tv, NULL);
  return (double)tv.tv_sec + (double)tv.tv_usec / 1000000.0;
}

void WhatCompileDefinitions() {
  std::cout << "Compiler flags:\n"
            << "          INTELQS_HAS_MPI --> [ NO]  (ranks are NCCL peers over NVLink)\n"
            << "                  USE_MKL --> [ NO]\n"
            << "                  _OPENMP --> [ NO]  (state loops are CUDA kernels, sm_100a)\n"
            << "            USE_MM_MALLOC --> [ NO]  (state lives in HBM)\n"
#ifdef NDEBUG
            << "                   NDEBUG --> [YES]\n";
#else
            << "                   NDEBUG --> [ NO]\n";
#endif
}

GateSpec1Q ConvertSpec2to1(GateSpec2Q spec) {
  switch (spec) {
    case GateSpec2Q::CHadamard: return GateSpec1Q::Hadamard;
    case GateSpec2Q::CRotationX: return GateSpec1Q::RotationX;
    case GateSpec2Q::CRotationY: return GateSpec1Q::RotationY;
    case GateSpec2Q::CRotationZ: return GateSpec1Q::RotationZ;
    case GateSpec2Q::CPauliX: return GateSpec1Q::PauliX;
    case GateSpec2Q::CPauliY: return GateSpec1Q::PauliY;
    case GateSpec2Q::CPauliZ: return GateSpec1Q::PauliZ;
    default: return GateSpec1Q::None;
  }
}

namespace mpi {

Environment *Environment::shared_instance = nullptr;
bool Environment::useful_rank = true;
bool Environment::is_verbose = true;

namespace {
iqsb_ctx *g_ctx = nullptr;
int g_rank = 0, g_size = 1, g_num_states = 1;
bool g_finalized = false;

[[noreturn]] void die(const char *what) {
  fprintf(stderr, "iqs (B200 engine): %s: %s\n", what, iqsb_last_error());
  throw std::runtime_error(std::string(what) + ": " + iqsb_last_error());
}

const char *env_first(const char *a, const char *b) {
  const char *v = getenv(a);
  if (v && *v) return v;
  v = getenv(b);
  return (v && *v) ? v : nullptr;
}

void create_context(int rank, int nranks, const void *uid, int device) {
  if (g_ctx) return;
  if (iqsb_init(rank, nranks, uid, device, &g_ctx) != IQSB_OK) die("cannot create the engine context");
  g_rank = rank;
  g_size = nranks;
}
}  // namespace

namespace detail {
namespace {
std::vector<LiveRegister> &Live() {
  static std::vector<LiveRegister> v;
  return v;
}
bool g_settling = false;
}  // namespace
void RegisterLive(const LiveRegister &r) { Live().push_back(r); }
void UnregisterLive(void *self) {
  auto &v = Live();
  for (std::size_t i = 0; i < v.size(); ++i)
    if (v[i].self == self) {
      v.erase(v.begin() + (std::ptrdiff_t)i);
      return;
    }
}
// creation order is the same on every rank (SPMD), so the collective steps pair up
void SettleAllLive() {
  if (g_settling) return;  // a register's own settle step may pass a barrier
  g_settling = true;
  try {
    std::vector<LiveRegister> snapshot = Live();
    for (auto &r : snapshot) r.settle(r.self);
  } catch (...) {
    g_settling = false;
    throw;
  }
  g_settling = false;
}
void ReleaseAllLive() {
  std::vector<LiveRegister> snapshot = Live();
  Live().clear();
  for (auto &r : snapshot) r.release(r.self);
}
}  // namespace detail

void Environment::Bootstrap() {
  if (g_ctx) return;
  const char *r = env_first("IQS_RANK", "RANK"), *s = env_first("IQS_NRANKS", "WORLD_SIZE");
  int rank = r ? atoi(r) : 0, world = s ? atoi(s) : 1;
  if (world <= 1) {
    create_context(0, 1, nullptr, -1);
    return;
  }
  // state ranks must be a power of two: the surplus ranks are "dummy" (reference mpi_env.cpp:239-300)
  int used = (int)floor_power_of_two(world);
  if (rank >= used) {
    useful_rank = false;
    return;
  }
  const char *lr = env_first("IQS_LOCAL_RANK", "LOCAL_RANK");
  int device = lr ? atoi(lr) : -1;
  std::string file;
  if (const char *f = getenv("IQS_UID_FILE")) file = f;
  else {
    const char *port = getenv("MASTER_PORT");
    file = std::string("/dev/shm/iqs_b200_uid_") + (port ? std::string(port) : toString((long)getppid()));
  }
  // Rendezvous file: the NCCL id followed by the wall-clock second rank 0 wrote it.  A file left behind
  // by a run that died (same MASTER_PORT, predictable name in a shared directory) is recognised by its
  // age and skipped instead of being used for a bootstrap that can never complete: rank 0 removes
  // whatever it finds before publishing, the others only accept a stamp from the last two minutes.
  unsigned char uid[IQSB_UNIQUE_ID_BYTES];
  const long long kMaxAgeSeconds = 120;
  if (rank == 0) {
    remove(file.c_str());
    if (iqsb_unique_id(uid) != IQSB_OK) die("cannot create the NCCL unique id");
    std::string tmp = file + ".tmp";
    // created exclusively, owner-only, never through a symlink somebody else planted in the shared directory
    remove(tmp.c_str());
    const int fd = open(tmp.c_str(), O_WRONLY | O_CREAT | O_EXCL | O_NOFOLLOW, 0600);
    FILE *fp = fd >= 0 ? fdopen(fd, "wb") : nullptr;
    long long stamp = (long long)time(nullptr);
    if (!fp || fwrite(uid, 1, sizeof(uid), fp) != sizeof(uid) || fwrite(&stamp, sizeof(stamp), 1, fp) != 1) throw std::runtime_error("cannot write " + tmp);
    fclose(fp);
    rename(tmp.c_str(), file.c_str());
  } else {
    bool have = false;
    for (int tries = 0; tries < 6000 && !have; ++tries) {
      if (FILE *fp = fopen(file.c_str(), "rb")) {
        long long stamp = 0;
        const bool complete = fread(uid, 1, sizeof(uid), fp) == sizeof(uid);
        const bool stamped = complete && fread(&stamp, sizeof(stamp), 1, fp) == 1;
        fclose(fp);
        // tools/iqsrun's per-launch file (IQS_UID_FILE, mkstemp name) carries no history: any complete id is good
        have = complete && (getenv("IQS_UID_FILE") != nullptr || (stamped && (long long)time(nullptr) - stamp <= kMaxAgeSeconds));
      }
      if (!have) std::this_thread::sleep_for(std::chrono::milliseconds(10));
    }
    if (!have) throw std::runtime_error("no fresh NCCL id in " + file + " after 60 s (is rank 0 running?)");
  }
  create_context(rank, used, uid, device);
  if (rank == 0 && !getenv("IQS_UID_FILE")) {
    iqsb_barrier(g_ctx);  // everybody is connected: the rendezvous file is no longer needed
    remove(file.c_str());
  } else if (!getenv("IQS_UID_FILE")) {
    iqsb_barrier(g_ctx);
  }
}

iqsb_ctx *Environment::Context() {
  if (!g_ctx) {
    if (g_finalized) throw std::runtime_error("iqs: the environment was finalized");
    Bootstrap();
    if (!g_ctx) throw std::runtime_error("iqs: this is a dummy rank (IsUsefulRank() == false); it must not touch registers");
  }
  return g_ctx;
}

void Environment::InitWithUniqueId(int rank, int nranks, const void *uid128, int device) {
  if (g_ctx) return;
  g_finalized = false;
  create_context(rank, nranks, uid128, device);
}
void Environment::GetUniqueId(void *out128) {
  if (iqsb_unique_id(out128) != IQSB_OK) die("cannot create the NCCL unique id");
}

Environment::Environment(int &, char **&, bool verbose) : inited_(true) {
  is_verbose = verbose;
  Bootstrap();
  shared_instance = this;
}
Environment::Environment() : inited_(true) {
  Bootstrap();
  shared_instance = this;
}
Environment::~Environment() {
  // like the reference's destructor (MPI_Finalize): the last act of a program.  Tearing the NCCL
  // communicator and the IPC mappings down here keeps process exit fast and orderly.
  if (shared_instance == this) {
    shared_instance = nullptr;
    if (g_ctx) {
      detail::ReleaseAllLive();  // registers that outlive the environment keep no handle into the freed context
      iqsb_finalize(g_ctx);
      g_ctx = nullptr;
      g_finalized = true;
    }
  }
}

void Environment::Init() {
  if (shared_instance != nullptr) throw std::runtime_error("iqs::mpi::Environment::Init: environment already initialized");
  g_finalized = false;
  shared_instance = new Environment();
}
void Environment::Init(int &argc, char **&argv) {
  if (shared_instance != nullptr) throw std::runtime_error("iqs::mpi::Environment::Init: environment already initialized");
  g_finalized = false;
  shared_instance = new Environment(argc, argv);
}
void Environment::Finalize() {
  if (shared_instance) {
    delete shared_instance;
    shared_instance = nullptr;
  }
  if (g_ctx) {
    detail::ReleaseAllLive();
    iqsb_finalize(g_ctx);
    g_ctx = nullptr;
    g_num_states = 1;
  }
  g_finalized = true;  // Context() no longer bootstraps silently; Init() starts a new environment
}

// Pool of states (reference src/mpi_env.cpp:279-360): the useful ranks are split into num_states
// groups, each holding one state.  Two splits are supported: one state over all GPUs (the default,
// the hot path) and one state PER GPU -- what the noisy-simulation tutorials and examples use to run
// an ensemble of trajectories in parallel.  Then every register is a single-GPU register (no
// global qubits, no exchange), and only IncoherentSumOverAllStatesOfPool crosses GPUs.
// Intermediate splits would need one NCCL communicator and peer table per group: not provided.
void Environment::UpdateStateComm(int new_num_states) {
  if (new_num_states == 1 || new_num_states == g_size) {
    if (g_ctx) iqsb_sync(g_ctx);
    g_num_states = new_num_states;
    return;
  }
  throw std::runtime_error("iqs::mpi::Environment::UpdateStateComm: supported pools are 1 state over all " + std::to_string(g_size) +
                           " GPUs or one state per GPU; got num_states = " + std::to_string(new_num_states));
}

int Environment::GetPoolRank() { return g_rank; }
int Environment::GetPoolSize() { return g_size; }
int Environment::GetStateRank() { return g_num_states == 1 ? g_rank : 0; }
int Environment::GetStateSize() { return g_num_states == 1 ? g_size : 1; }
int Environment::GetNumRanksPerNode() { return g_size; }
int Environment::GetNumNodes() { return 1; }
int Environment::GetNodeId() { return 0; }
int Environment::GetStateId() { return g_num_states == 1 ? 0 : g_rank; }
int Environment::GetNumStates() { return g_num_states; }
// The reference renumbers ranks to implement X/Y on a global qubit without data movement
// (spec-v1 only, flagged buggy: qureg_apply1qubitgate.cpp:61-99).  Not used by this engine.
void Environment::RemapStateRank(int) {}

// sum over the pool divided by the ranks per state (mpi_env.cpp:467-486): every rank of a state
// holds the same value, so this is the sum over states
template <class Type>
Type Environment::IncoherentSumOverAllStatesOfPool(Type local_value) {
  if (g_size == 1) return local_value;
  double v = (double)local_value;
  if (iqsb_allreduce_f64(Environment::Context(), &v, 1, IQSB_SUM) != IQSB_OK) die("allreduce over the pool failed");
  return (Type)(v / double(GetStateSize()));
}
template float Environment::IncoherentSumOverAllStatesOfPool<float>(float);
template double Environment::IncoherentSumOverAllStatesOfPool<double>(double);

// Barriers are the program's synchronisation points: they run every live register's queued gates,
// put its amplitudes back in the reference's order (a collective step when qubits were moved between
// local and rank bits) and drain the engine's stream -- after StateBarrier() a rank may read its shard.
void PoolBarrier() {
  if (g_ctx) detail::SettleAllLive();
  if (g_ctx && iqsb_barrier(g_ctx) != IQSB_OK) die("barrier failed");
}
void StateBarrier() {
  if (Environment::GetStateSize() > 1) PoolBarrier();
  else {
    if (g_ctx) detail::SettleAllLive();
    if (g_ctx && iqsb_sync(g_ctx) != IQSB_OK) die("synchronisation failed");
  }
}
void Barrier() { PoolBarrier(); }

namespace {
void PrintInOrder(const std::string &s, bool all, int rank, int size, void (*barrier)()) {
  if (all) {
    for (int r = 0; r < size; ++r) {
      if (r == rank) {
        printf("[|%d>:%3d] %s\n", Environment::GetStateId(), rank, s.c_str());
        fflush(stdout);
      }
      barrier();
    }
  } else if (rank == 0) {
    std::cout << s << std::endl;
  }
}
}  // namespace
void StatePrint(std::string s, bool all) { PrintInOrder(s, all, Environment::GetStateRank(), Environment::GetStateSize(), StateBarrier); }
void PoolPrint(std::string s, bool all) { PrintInOrder(s, all, Environment::GetPoolRank(), Environment::GetPoolSize(), PoolBarrier); }
void Print(std::string s, bool all) { PoolPrint(s, all); }

// collectives of ONE state: nothing to do when the state lives on one GPU
void AllreduceDouble(double *inout, int n, ReduceOp op) {
  if (Environment::GetStateSize() == 1) return;
  if (iqsb_allreduce_f64(Environment::Context(), inout, n, op == MAX ? IQSB_MAX : IQSB_SUM) != IQSB_OK) die("allreduce failed");
}
void BcastDouble(double *inout, int n, int root) {
  if (Environment::GetStateSize() == 1) return;
  if (iqsb_bcast_f64(Environment::Context(), inout, n, root) != IQSB_OK) die("broadcast failed");
}

}  // namespace mpi
}  // namespace iqs

#include "../include/qureg_version.hpp"
namespace iqs {
// the QASM interface prints this string (interface/src/interface_api_version.cpp:16-21); it names the API level
std::string GetQhipsterVersion() { return QHIPSTER_VERSION_STRING; }
}  // namespace iqs
