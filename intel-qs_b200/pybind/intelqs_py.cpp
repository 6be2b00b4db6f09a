// intelqs_py.cpp -- Python module `intelqs_py` over the B200 iqs::QubitRegister<ComplexDP>.
//
// Same Python surface as the reference's module for the hot path (reference
// pybind11/intelqs_py.cpp:56-426: EnvInit/EnvFinalize, QubitRegister with the NumPy buffer
// protocol, named gates, custom 2x2 gates from a complex128 array, measurement, expectation
// values, RandomNumberGenerator, MPIEnvironment statics), so notebooks and scripts written for
// Intel-QS run unchanged.  Additions (marked "B200") expose what the reference's module lacks.
#include <pybind11/complex.h>
#include <pybind11/iostream.h>
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <string>
#include <vector>

#include "iqsb.h"
#include "qaoa_features.hpp"
#include "qureg.hpp"

namespace py = pybind11;
using Environment = iqs::mpi::Environment;
using Reg = iqs::QubitRegister<ComplexDP>;

namespace {

// chi matrices of 1- and 2-qubit channels (reference intelqs_py.cpp:94-163): chi[i, j] access,
// SolveEigenSystem, Print; the B200 module also reads the eigensystem back
template <unsigned N>
void BindChi(py::module &m, const char *name, const char *repr) {
  using Chi = iqs::ChiMatrix<ComplexDP, N, 32>;
  py::class_<Chi>(m, name)
      .def(py::init<>())
      .def("__getitem__", [](const Chi &a, std::pair<py::ssize_t, py::ssize_t> i) {
        if (i.first < 0 || i.first >= (py::ssize_t)N || i.second < 0 || i.second >= (py::ssize_t)N) throw py::index_error();
        return a(i.first, i.second);
      }, py::is_operator())
      .def("__setitem__", [](Chi &a, std::pair<py::ssize_t, py::ssize_t> i, ComplexDP value) {
        if (i.first < 0 || i.first >= (py::ssize_t)N || i.second < 0 || i.second >= (py::ssize_t)N) throw py::index_error();
        a(i.first, i.second) = value;
      }, py::is_operator())
      .def("SolveEigenSystem", &Chi::SolveEigenSystem)
      .def("EigensystemOfIdealHadamardChannel", &Chi::EigensystemOfIdealHadamardChannel)
      .def("GetEigenValues", &Chi::GetEigenValues, "B200: eigenvalues (ascending)")
      .def("GetEigenProbabilities", &Chi::GetEigenProbabilities, "B200: |E_k| / sum |E_k|")
      .def("GetEigenVectors", &Chi::GetEigenVectors, "B200: standardised, renormalised eigenvectors in the Pauli basis")
      .def("Print", [](Chi &a, bool with_eigensystem) {
        py::scoped_ostream_redirect stream(std::cout, py::module::import("sys").attr("stdout"));
        a.Print(with_eigensystem);
      }, py::arg("with_eigensystem") = true)
      .def("__repr__", [repr](const Chi &) { return std::string(repr); });
}

TM2x2<ComplexDP> FromArray(py::array_t<ComplexDP, py::array::c_style | py::array::forcecast> matrix) {
  py::buffer_info buf = matrix.request();
  if (buf.ndim != 2) throw std::runtime_error("Number of dimensions must be two.");
  if (buf.shape[0] != 2 || buf.shape[1] != 2) throw std::runtime_error("Input shape is not 2x2.");
  ComplexDP *ptr = (ComplexDP *)buf.ptr;
  TM2x2<ComplexDP> m;
  m(0, 0) = ptr[0];
  m(0, 1) = ptr[1];
  m(1, 0) = ptr[2];
  m(1, 1) = ptr[3];
  return m;
}

TM4x4<ComplexDP> FromArray4(py::array_t<ComplexDP, py::array::c_style | py::array::forcecast> matrix) {
  py::buffer_info buf = matrix.request();
  if (buf.ndim != 2 || buf.shape[0] != 4 || buf.shape[1] != 4) throw std::runtime_error("Input shape is not 4x4.");
  ComplexDP *ptr = (ComplexDP *)buf.ptr;
  TM4x4<ComplexDP> m;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) m(i, j) = ptr[4 * i + j];
  return m;
}

void EnvInit() { Environment::Init(); }
void EnvFinalize() { Environment::Finalize(); }
void EnvFinalizeDummyRanks() {
  if (Environment::GetSharedInstance() && Environment::GetSharedInstance()->IsUsefulRank() == false) Environment::Finalize();
}

}  // namespace

PYBIND11_MODULE(intelqs_py, m) {
  m.doc() = "pybind11 wrap of the B200-native Intel Quantum Simulator engine";

  m.def("EnvInit", &EnvInit, "Initialize the rank environment (NCCL bootstrap when several ranks are launched)");
  m.def("EnvFinalize", &EnvFinalize, "Finalize the rank environment");
  m.def("EnvFinalizeDummyRanks", &EnvFinalizeDummyRanks, "Finalize the dummy ranks of the environment");
  // B200: explicit bootstrap for launchers that move the NCCL id themselves (torch.distributed)
  m.def("EnvUniqueId", []() {
    std::string id(IQSB_UNIQUE_ID_BYTES, '\0');
    Environment::GetUniqueId(&id[0]);
    return py::bytes(id);
  });
  m.def("EnvInitWithUniqueId", [](int rank, int nranks, py::bytes uid, int device) {
    std::string s = uid;
    if (nranks > 1 && s.size() != IQSB_UNIQUE_ID_BYTES) throw std::runtime_error("the unique id must be 128 bytes");
    Environment::InitWithUniqueId(rank, nranks, nranks > 1 ? s.data() : nullptr, device);
  }, py::arg("rank"), py::arg("nranks"), py::arg("uid"), py::arg("device") = -1);
  m.def("SetContractedArithmetic", [](bool fma) {
    if (iqsb_set_arith(Environment::Context(), fma ? IQSB_ARITH_FMA : IQSB_ARITH_EXACT) != IQSB_OK) throw std::runtime_error(iqsb_last_error());
  }, "B200: False (default) = exact arithmetic in the reference's operation order; True = contracted multiply-adds in the fused kernel");
  m.def("LaunchCount", []() { return (unsigned long long)iqsb_launch_count(Environment::Context()); }, "B200: kernels launched so far");
  m.def("NvlinkBytes", []() { return (unsigned long long)iqsb_nvlink_bytes(Environment::Context()); }, "B200: bytes moved over NVLink so far");
  m.def("DeviceTimerStart", []() { iqsb_timer_start(Environment::Context()); }, "B200: CUDA event on the engine's stream");
  m.def("DeviceTimerStop", []() {
    double ms = 0;
    iqsb_timer_stop(Environment::Context(), &ms);
    return ms;
  }, "B200: milliseconds since DeviceTimerStart, measured on the device");
  m.def("DeviceSync", []() {
    if (iqsb_sync(Environment::Context()) != IQSB_OK) throw std::runtime_error(iqsb_last_error());
  }, "B200: wait for the engine's stream (no gate queue is flushed, no placement restored)");
  m.def("DeviceProfile", [](bool on) {
    if (iqsb_profile(Environment::Context(), on ? 1 : 0) != IQSB_OK) throw std::runtime_error(iqsb_last_error());
  }, "B200: start (clearing) / stop per-kernel-class timing with CUDA events on the engine's stream");
  m.def("DeviceProfileRead", []() {
    std::vector<char> buf(1 << 16);
    if (iqsb_profile_read(Environment::Context(), buf.data(), buf.size()) != IQSB_OK) throw std::runtime_error(iqsb_last_error());
    return std::string(buf.data());
  }, "B200: JSON text {overflow, classes: [{name, launches, ms, bytes}]} of the profiled region");

  py::class_<iqs::RandomNumberGenerator<double>>(m, "RandomNumberGenerator")
      .def(py::init<>())
      .def("GetSeed", &iqs::RandomNumberGenerator<double>::GetSeed)
      .def("SetSeedStreamPtrs", &iqs::RandomNumberGenerator<double>::SetSeedStreamPtrs)
      .def("SkipeAhead", &iqs::RandomNumberGenerator<double>::SkipAhead)
      .def("SkipAhead", &iqs::RandomNumberGenerator<double>::SkipAhead)
      .def("GetUniformRandomNumbers",
           [](iqs::RandomNumberGenerator<double> &rng, std::size_t size, double a, double b, std::string shared) {
             std::vector<double> v(size);
             rng.UniformRandomNumbers(v.data(), size, a, b, shared);
             return v;
           },
           "Return an array of 'size' random number from the uniform distribution [a,b[.")
      .def("__repr__", [](const iqs::RandomNumberGenerator<double> &) { return "<RandomNumberGenerator (std::mt19937 streams)>"; });

  BindChi<4>(m, "CM4x4", "<ChiMatrix for 1-qubit channel>");
  BindChi<16>(m, "CM16x16", "<ChiMatrix for 2-qubit channel>");

  py::class_<Reg>(m, "QubitRegister", py::buffer_protocol(), py::dynamic_attr())
      .def(py::init<>())
      .def(py::init<const Reg &>())
      .def(py::init<std::size_t, std::string, std::size_t, std::size_t>())
      .def("NumQubits", &Reg::NumQubits)
      .def("GlobalSize", &Reg::GlobalSize)
      .def("LocalSize", &Reg::LocalSize)
      .def("__getitem__", [](const Reg &a, std::size_t index) {
        if (index >= a.LocalSize()) throw py::index_error();
        return a[index];
      }, py::is_operator())
      .def("__setitem__", [](Reg &a, std::size_t index, ComplexDP value) {
        if (index >= a.LocalSize()) throw py::index_error();
        a[index] = value;
      }, py::is_operator())
      // NumPy view of the local shard (zero-copy on one rank: the shard is managed memory)
      .def_buffer([](Reg &reg) -> py::buffer_info {
        return py::buffer_info(reg.RawState(), sizeof(ComplexDP), py::format_descriptor<ComplexDP>::format(), 1,
                               {reg.LocalSize()}, {sizeof(ComplexDP)});
      })
      // one-qubit gates
      .def("ApplyRotationX", &Reg::ApplyRotationX)
      .def("ApplyRotationY", &Reg::ApplyRotationY)
      .def("ApplyRotationZ", &Reg::ApplyRotationZ)
      .def("ApplyPauliX", &Reg::ApplyPauliX)
      .def("ApplyPauliY", &Reg::ApplyPauliY)
      .def("ApplyPauliZ", &Reg::ApplyPauliZ)
      .def("ApplyPauliSqrtX", &Reg::ApplyPauliSqrtX)
      .def("ApplyPauliSqrtY", &Reg::ApplyPauliSqrtY)
      .def("ApplyPauliSqrtZ", &Reg::ApplyPauliSqrtZ)
      .def("ApplyT", &Reg::ApplyT)
      .def("ApplyRotationXY", &Reg::ApplyRotationXY)
      .def("ApplyHadamard", &Reg::ApplyHadamard)
      // two-qubit gates
      .def("ApplySwap", &Reg::ApplySwap)
      .def("ApplyCRotationX", &Reg::ApplyCRotationX)
      .def("ApplyCRotationY", &Reg::ApplyCRotationY)
      .def("ApplyCRotationZ", &Reg::ApplyCRotationZ)
      .def("ApplyCPauliX", &Reg::ApplyCPauliX)
      .def("ApplyCPauliY", &Reg::ApplyCPauliY)
      .def("ApplyCPauliZ", &Reg::ApplyCPauliZ)
      .def("ApplyCPauliSqrtZ", &Reg::ApplyCPauliSqrtZ)
      .def("ApplyCHadamard", &Reg::ApplyCHadamard)
      .def("Apply1QubitGate", [](Reg &a, unsigned qubit, py::array_t<ComplexDP, py::array::c_style | py::array::forcecast> matrix) {
        a.Apply1QubitGate(qubit, FromArray(matrix));
      }, "Apply custom 1-qubit gate.")
      .def("ApplyControlled1QubitGate", [](Reg &a, unsigned control, unsigned qubit, py::array_t<ComplexDP, py::array::c_style | py::array::forcecast> matrix) {
        a.ApplyControlled1QubitGate(control, qubit, FromArray(matrix));
      }, "Apply custom controlled-1-qubit gate.")
      .def("ApplyToffoli", &Reg::ApplyToffoli)
      .def("GetOverallSignOfChannels", &Reg::GetOverallSignOfChannels)
      .def("ApplyChannel", [](Reg &a, unsigned qubit, iqs::ChiMatrix<ComplexDP, 4, 32> chi) { a.ApplyChannel(qubit, chi); },
           "Apply 1-qubit channel provided via its chi-matrix.")
      .def("ApplyChannel", [](Reg &a, unsigned qubit1, unsigned qubit2, iqs::ChiMatrix<ComplexDP, 16, 32> chi) { a.ApplyChannel(qubit1, qubit2, chi); },
           "Apply 2-qubit channel provided via its chi-matrix.")
      // state initialization
      .def("Initialize", (void (Reg::*)(std::string, std::size_t)) & Reg::Initialize)
      .def("TurnOnSpecialize", &Reg::TurnOnSpecialize)
      .def("TurnOffSpecialize", &Reg::TurnOffSpecialize)
      .def("TurnOnSpecializeV2", &Reg::TurnOnSpecializeV2)
      .def("TurnOffSpecializeV2", &Reg::TurnOffSpecializeV2)
      .def("ResetRngPtr", &Reg::ResetRngPtr)
      .def("SetRngPtr", &Reg::SetRngPtr)
      .def("SetSeedRngPtr", &Reg::SetSeedRngPtr)
      // measurement
      .def("GetProbability", &Reg::GetProbability)
      .def("CollapseQubit", &Reg::CollapseQubit)
      .def("Normalize", &Reg::Normalize)
      .def("AmplitudeWiseScalarMultiplication", &Reg::AmplitudeWiseScalarMultiplication)
      .def("ExpectationValue", &Reg::ExpectationValue)
      .def("ComputeNorm", &Reg::ComputeNorm)
      .def("ComputeOverlap", &Reg::ComputeOverlap)
      // noise gates (clients of Apply1QubitGate)
      .def("GetT1", &Reg::GetT1)
      .def("GetT2", &Reg::GetT2)
      .def("GetTphi", &Reg::GetTphi)
      .def("SetNoiseTimescales", &Reg::SetNoiseTimescales)
      .def("ApplyNoiseGate", &Reg::ApplyNoiseGate)
      .def("Print", [](Reg &a, std::string description) {
        py::scoped_ostream_redirect stream(std::cout, py::module::import("sys").attr("stdout"));
        a.Print(description, {});
      }, "Print the quantum state with an initial description.")
      // ---- B200: methods of the C++ class that the reference's module does not bind ----------
      .def("ApplyCPhaseRotation", &Reg::ApplyCPhaseRotation)
      .def("ApplyISwap", &Reg::ApplyISwap)
      .def("ApplySqrtISwap", &Reg::ApplySqrtISwap)
      .def("Apply4thRootISwap", &Reg::Apply4thRootISwap)
      .def("ApplyISwapRotation", [](Reg &a, unsigned q1, unsigned q2, py::array_t<ComplexDP, py::array::c_style | py::array::forcecast> matrix) {
        a.ApplyISwapRotation(q1, q2, FromArray(matrix));
      })
      .def("ApplyDiag", [](Reg &a, unsigned q1, unsigned q2, py::array_t<ComplexDP, py::array::c_style | py::array::forcecast> matrix) {
        a.ApplyDiag(q1, q2, FromArray4(matrix));
      })
      .def("Apply2QubitGate", [](Reg &a, unsigned qh, unsigned ql, py::array_t<ComplexDP, py::array::c_style | py::array::forcecast> matrix) {
        a.Apply2QubitGate(qh, ql, FromArray4(matrix));
      })
      .def("PermuteQubits", &Reg::PermuteQubits, py::arg("new_map"), py::arg("style_of_map") = "direct")
      .def("EmulateSwap", &Reg::EmulateSwap)
      .def("GetQubitMap", [](Reg &a) { return a.qubit_permutation->map; })
      .def("TurnOnFusion", &Reg::TurnOnFusion, py::arg("log2llc") = 20)
      .def("TurnOffFusion", &Reg::TurnOffFusion)
      .def("IsFusionEnabled", &Reg::IsFusionEnabled)
      .def("ApplyFusedGates", &Reg::ApplyFusedGates)
      .def("GetGlobalAmplitude", &Reg::GetGlobalAmplitude)
      .def("SetGlobalAmplitude", &Reg::SetGlobalAmplitude)
      .def("MaxAbsDiff", [](Reg &a, Reg &b) { return a.MaxAbsDiff(b); })
      .def("MaxL2NormDiff", &Reg::MaxL2NormDiff)
      .def("IsClassicalBit", &Reg::IsClassicalBit, py::arg("qubit"), py::arg("tolerance") = 1.e-13)
      .def("GetClassicalValue", &Reg::GetClassicalValue, py::arg("qubit"), py::arg("tolerance") = 1.e-13)
      .def("ExpectationValueX", &Reg::ExpectationValueX, py::arg("qubit"), py::arg("coeff") = 1.)
      .def("ExpectationValueY", &Reg::ExpectationValueY, py::arg("qubit"), py::arg("coeff") = 1.)
      .def("ExpectationValueZ", &Reg::ExpectationValueZ, py::arg("qubit"), py::arg("coeff") = 1.)
      .def("Entropy", &Reg::Entropy)
      .def("GoogleStats", &Reg::GoogleStats)
      .def("EnableStatistics", &Reg::EnableStatistics)
      .def("GetStatistics", &Reg::GetStatistics)
      .def("DisableStatistics", &Reg::DisableStatistics)
      .def("AmplitudeWiseSum", [](Reg &a, Reg &b, ComplexDP f) { a.AmplitudeWiseSum(b, f); }, py::arg("psi"), py::arg("factor") = ComplexDP(1, 0))
      .def("SyncToHost", &Reg::SyncToHost)
      .def("Upload", [](Reg &a, py::array_t<ComplexDP, py::array::c_style | py::array::forcecast> v) {
        py::buffer_info buf = v.request();
        if ((std::size_t)buf.size != a.LocalSize()) throw std::runtime_error("Upload: expected LocalSize() amplitudes");
        a.SyncToHost();
        if (iqsb_upload(a.DeviceState(), buf.ptr, 0, a.LocalSize()) != IQSB_OK) throw std::runtime_error(iqsb_last_error());
      }, "B200: replace the local shard by a complex128 array (one host-to-device copy)")
      .def("Download", [](Reg &a) {
        a.SyncToHost();
        py::array_t<ComplexDP> out(a.LocalSize());
        if (iqsb_download(a.DeviceState(), out.mutable_data(), 0, a.LocalSize()) != IQSB_OK) throw std::runtime_error(iqsb_last_error());
        return out;
      }, "B200: copy of the local shard as a complex128 array (one device-to-host copy)");

  // QAOA helpers (reference pybind11/intelqs_py.cpp:362-393): the loops run on the device
  m.def("InitializeVectorAsMaxCutCostFunction", &iqs::qaoa::InitializeVectorAsMaxCutCostFunction<ComplexDP>,
        "Use IQS vector to store a large real vector and not as a quantum state.");
  m.def("InitializeVectorAsWeightedMaxCutCostFunction", &iqs::qaoa::InitializeVectorAsWeightedMaxCutCostFunction<ComplexDP>,
        "Use IQS vector to store a large real vector and not as a quantum state.");
  m.def("ImplementQaoaLayerBasedOnCostFunction", &iqs::qaoa::ImplementQaoaLayerBasedOnCostFunction<ComplexDP>, "Implement exp(-i gamma C)|psi>.");
  m.def("GetExpectationValueFromCostFunction", &iqs::qaoa::GetExpectationValueFromCostFunction<ComplexDP>,
        "Get expectation value from the cost function.");
  m.def("GetExpectationValueSquaredFromCostFunction", &iqs::qaoa::GetExpectationValueSquaredFromCostFunction<ComplexDP>,
        "Get expectation value squared from the cost function.");
  m.def("GetHistogramFromCostFunction", &iqs::qaoa::GetHistogramFromCostFunction<ComplexDP>, "Get histogram instead of just the expectation value.");
  m.def("GetHistogramFromCostFunctionWithWeightsRounded", &iqs::qaoa::GetHistogramFromCostFunctionWithWeightsRounded<ComplexDP>,
        "Get histogram instead of just the expectation value for a weighted graph, with all cut values rounded down.");
  m.def("GetHistogramFromCostFunctionWithWeightsBinned", &iqs::qaoa::GetHistogramFromCostFunctionWithWeightsBinned<ComplexDP>,
        "Get histogram instead of just the expectation value for a weighted graph, with specified bin width.");

  py::class_<Environment>(m, "MPIEnvironment")
      .def(py::init<>())
      .def_static("GetRank", &Environment::GetRank)
      .def_static("IsUsefulRank", &Environment::IsUsefulRank)
      .def_static("GetSizeWorldComm", []() { return Environment::GetPoolSize(); }, "Number of processes when the environment was first created.")
      .def_static("GetPoolRank", &Environment::GetPoolRank)
      .def_static("GetStateRank", &Environment::GetStateRank)
      .def_static("GetPoolSize", &Environment::GetPoolSize)
      .def_static("GetStateSize", &Environment::GetStateSize)
      .def_static("GetNumRanksPerNode", &Environment::GetNumRanksPerNode)
      .def_static("GetNumNodes", &Environment::GetNumNodes)
      .def_static("GetStateId", &Environment::GetStateId)
      .def_static("GetNumStates", &Environment::GetNumStates)
      .def_static("Barrier", &iqs::mpi::Barrier)
      .def_static("PoolBarrier", &iqs::mpi::PoolBarrier)
      .def_static("StateBarrier", &iqs::mpi::StateBarrier)
      .def_static("IncoherentSumOverAllStatesOfPool", &Environment::IncoherentSumOverAllStatesOfPool<double>)
      .def_static("UpdateStateComm", &Environment::UpdateStateComm);
}
