// QubitRegisterMetric.hpp -- a QubitRegister that also counts gates and a greedy circuit depth
// (interface of reference include/QubitRegisterMetric.hpp:10-37).  A thin wrapper: every gate is
// forwarded to the B200 engine.
#ifndef QUBIT_REGISTER_METRIC_HPP
#define QUBIT_REGISTER_METRIC_HPP

#include <algorithm>
#include <vector>

// The reference's header injects namespace std into every translation unit that includes
// qureg.hpp, and some of its own programs rely on that (benchmarks/basic_code_for_scaling.cpp:156
// uses an unqualified `ofstream`).  Kept for source compatibility.
using namespace std;

namespace iqs {

template <class Type = ComplexDP>
class QubitRegisterMetric : public QubitRegister<Type> {
  int iTotalQubitGateCount = 0;
  int iOneQubitGateCount = 0;
  int iTwoQubitGateCount = 0;
  std::vector<int> vParallelDepth;
  void OneQubitIncrements(int q) {
    ++iTotalQubitGateCount;
    ++iOneQubitGateCount;
    ++vParallelDepth[q];
  }
  void TwoQubitIncrements(int q1, int q2) {
    ++iTotalQubitGateCount;
    ++iTwoQubitGateCount;
    int depth = std::max(vParallelDepth[q1], vParallelDepth[q2]) + 1;
    vParallelDepth[q1] = vParallelDepth[q2] = depth;
  }

 public:
  QubitRegisterMetric(int iNQubits) : QubitRegister<Type>(iNQubits), vParallelDepth(iNQubits, 0) {}

  int GetTotalQubitGateCount() { return iTotalQubitGateCount; }
  int GetOneQubitGateCount() { return iOneQubitGateCount; }
  int GetTwoQubitGateCount() { return iTwoQubitGateCount; }
  int GetParallelDepth() { return *std::max_element(vParallelDepth.begin(), vParallelDepth.end()); }

  void ApplyHadamard(int q) {
    QubitRegister<Type>::ApplyHadamard(q);
    OneQubitIncrements(q);
  }
  void ApplyRotationX(int q, double theta) {
    QubitRegister<Type>::ApplyRotationX(q, theta);
    OneQubitIncrements(q);
  }
  void ApplyRotationY(int q, double theta) {
    QubitRegister<Type>::ApplyRotationY(q, theta);
    OneQubitIncrements(q);
  }
  void ApplyRotationZ(int q, double theta) {
    QubitRegister<Type>::ApplyRotationZ(q, theta);
    OneQubitIncrements(q);
  }
  void ApplyCPauliX(int q1, int q2) {
    QubitRegister<Type>::ApplyCPauliX(q1, q2);
    TwoQubitIncrements(q1, q2);
  }
  void ApplyControlled1QubitGate(int q1, int q2, iqs::TinyMatrix<Type, 2, 2, 32> V) {
    QubitRegister<Type>::ApplyControlled1QubitGate(q1, q2, V);
    TwoQubitIncrements(q1, q2);
  }
};

}  // namespace iqs
#endif
