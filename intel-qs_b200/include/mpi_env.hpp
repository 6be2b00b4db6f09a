// mpi_env.hpp -- the rank environment of the B200 engine.
//
// Interface of reference include/mpi_env.hpp:46-145 (class iqs::mpi::Environment with the same
// static queries, Init/Finalize, barriers and Print helpers).  There is no MPI underneath: one
// process drives one GPU, the ranks of a job form an NCCL communicator over NVLink/NVSwitch,
// and "state rank r" owns global amplitudes [r*L, (r+1)*L) exactly as in the reference
// (src/qureg_init.cpp:97-107).
//
// Rank bootstrap (there is no mpirun in the image): Init reads
//   IQS_RANK | RANK,  IQS_NRANKS | WORLD_SIZE,  IQS_LOCAL_RANK | LOCAL_RANK   (torchrun sets the latter set)
// and exchanges the 128-byte NCCL id through the file named by IQS_UID_FILE
// (default /dev/shm/iqs_b200_uid_<MASTER_PORT|ppid>).  `tools/iqsrun -n N prog args` sets all of it.
// A host program that already has a transport can call InitWithUniqueId instead.
#ifndef IQS_MPI_ENV_HPP
#define IQS_MPI_ENV_HPP

#include <stdexcept>
#include <string>
#include <unistd.h>

struct iqsb_ctx;

namespace iqs {
namespace mpi {

class Environment {
 public:
  Environment(int &argc, char **&argv, bool is_verbose = true);
  Environment();
  ~Environment();
  Environment(Environment const &) = delete;
  Environment &operator=(Environment const &) = delete;

  // Pool-of-states is "replicas only" on this engine: num_states must be 1.
  static void UpdateStateComm(int num_states);

  static bool IsUsefulRank() { return useful_rank; }

  static int GetPoolRank();
  static int GetStateRank();
  static int GetRank() { return GetStateRank(); }
  static int GetPoolSize();
  static int GetStateSize();
  static int GetSize() { return GetStateSize(); }

  template <class Type>
  static Type IncoherentSumOverAllStatesOfPool(Type local_value);

  static int GetNumRanksPerNode();
  static int GetNumNodes();
  static int GetNodeId();
  static int GetStateId();
  static int GetNumStates();
  static void RemapStateRank(int newme);

  static void Init();
  static void Init(int &argc, char **&argv);
  static void Finalize();
  static Environment *GetSharedInstance() { return shared_instance; }

  // ---- B200 extensions -------------------------------------------------------------------
  // the engine context of this process (created on first use: rank 0 of 1 unless Init ran)
  static iqsb_ctx *Context();
  // explicit bootstrap for hosts that move the NCCL id themselves (e.g. torch.distributed)
  static void InitWithUniqueId(int rank, int nranks, const void *uid128, int device = -1);
  static void GetUniqueId(void *out128);

 private:
  static void Bootstrap();
  static Environment *shared_instance;
  bool inited_;
  static bool useful_rank;
  static bool is_verbose;
};

void PoolBarrier();
void StateBarrier();
void Barrier();

void PoolPrint(std::string s, bool all = false);
void StatePrint(std::string s, bool all = false);
void Print(std::string s, bool all = false);

}  // namespace mpi
}  // namespace iqs
#endif
