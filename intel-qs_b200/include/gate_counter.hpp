// gate_counter.hpp -- counts gates per program qubit and a greedy depth (interface of reference
// include/gate_counter.hpp:18-95); host only.
#ifndef IQS_GATE_COUNTER_HPP
#define IQS_GATE_COUNTER_HPP
#include <algorithm>
#include <cstdio>
#include <vector>

#include "mpi_env.hpp"

namespace iqs {
class GateCounter {
 private:
  int num_qubits;
  int total_gate_count = 0;
  int one_qubit_gate_count = 0;
  int two_qubit_gate_count = 0;
  std::vector<int> parallel_depth;

 public:
  GateCounter(int new_num_qubits) : num_qubits(new_num_qubits), parallel_depth(new_num_qubits, 0) {}
  ~GateCounter() {}
  void Reset() { parallel_depth.assign(num_qubits, 0); }
  int GetTotalGateCount() { return total_gate_count; }
  int GetOneQubitGateCount() { return one_qubit_gate_count; }
  int GetTwoQubitGateCount() { return two_qubit_gate_count; }
  int GetParallelDepth() { return parallel_depth.empty() ? 0 : *std::max_element(parallel_depth.begin(), parallel_depth.end()); }
  void OneQubitIncrement(int qubit) {
    ++total_gate_count;
    ++one_qubit_gate_count;
    ++parallel_depth[qubit];
  }
  void TwoQubitIncrement(int qubit_0, int qubit_1) {
    ++total_gate_count;
    ++two_qubit_gate_count;
    int d = std::max(parallel_depth[qubit_0], parallel_depth[qubit_1]) + 1;
    parallel_depth[qubit_0] = parallel_depth[qubit_1] = d;
  }
  void Breakdown() {
    if (iqs::mpi::Environment::GetStateRank() == 0)
      printf("The quantum circuit is composed of %d one-qubit gates and %d two-qubitgates, for a total of %d gates.\n"
             "The greedy depth (all gates lasting one clock cycle) is %d.\n",
             GetOneQubitGateCount(), GetTwoQubitGateCount(), GetTotalGateCount(), GetParallelDepth());
  }
};
}  // namespace iqs
#endif
