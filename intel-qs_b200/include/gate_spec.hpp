// gate_spec.hpp -- gate tags of "specialization v2" (interface of reference include/gate_spec.hpp:6-27).
// On the GPU the tags are hints only: every gate runs the same HBM-bound kernel.
#ifndef GATE_SPEC_HPP
#define GATE_SPEC_HPP
namespace iqs {
enum class GateSpec1Q { Hadamard = 0, RotationX, RotationY, RotationZ, PauliX, PauliY, PauliZ, T, None };
enum class GateSpec2Q { CHadamard = 0, CRotationX, CRotationY, CRotationZ, CPauliX, CPauliY, CPauliZ, CPhase, None };
GateSpec1Q ConvertSpec2to1(GateSpec2Q spec);
}  // namespace iqs
#endif
