// mpi_utils.hpp -- scalar collectives of the state communicator.
// The reference wraps MPI_Allreduce / MPI_Bcast / MPI_Sendrecv (include/mpi_utils.hpp:34-78); here the
// same roles are played by NCCL on a tiny device buffer (C ABI: iqsb_allreduce_f64 / iqsb_bcast_f64).
#ifndef IQS_MPI_UTILS_HPP
#define IQS_MPI_UTILS_HPP
#include <cstddef>
namespace iqs {
namespace mpi {
enum ReduceOp { SUM = 0, MAX = 1 };
// in-place all-reduce of n doubles over the ranks of the state (no-op for one rank)
void AllreduceDouble(double *inout, int n, ReduceOp op);
// broadcast n doubles from state rank `root`
void BcastDouble(double *inout, int n, int root);

// Registry of the live registers of this process (B200 engine).  The barriers of this namespace are
// the program's explicit synchronisation points: they run every register's pending gates and put
// its amplitudes back in the reference's order, so that after StateBarrier() a rank may read its
// shard on its own.  Environment::Finalize() releases the device memory of registers still alive.
namespace detail {
struct LiveRegister {
  void *self;
  void (*settle)(void *self);   // run queued gates + restore the reference's amplitude order (collective)
  void (*release)(void *self);  // free device memory (the context is going away)
};
void RegisterLive(const LiveRegister &r);
void UnregisterLive(void *self);
void SettleAllLive();
void ReleaseAllLive();
}  // namespace detail
}  // namespace mpi
}  // namespace iqs
#endif
