"""Python face of the drop-in (`intelqs_py`), with the reference's own unit tests restated
(unit_test/include/*.hpp) and its README golden vector."""
import math
import os
import sys

import numpy as np
import pytest

from pkg import circuits as C

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ACC = 1e-15  # accepted_error_ of the reference fixtures


@pytest.fixture(scope="module")
def iqs():
    sys.path.insert(0, os.path.join(ROOT, "intel-qs_b200", "lib"))
    import intelqs_py

    return intelqs_py


def amps(psi):
    return np.array(psi, copy=True)


def test_readme_golden_vector(iqs):
    """notebooks/print_distributed_state.py:50-53: 4 qubits, H on all, RZ(3, pi/3)."""
    psi = iqs.QubitRegister(4, "base", 0, 0)
    for q in range(4):
        psi.ApplyHadamard(q)
    psi.ApplyRotationZ(3, math.pi / 3)
    a = amps(psi)
    assert np.allclose(a[:8], 0.21650635 - 0.125j, atol=1e-8)
    assert np.allclose(a[8:], 0.21650635 + 0.125j, atol=1e-8)


def test_numpy_view_is_zero_copy_and_writable(iqs):
    """pybind11/intelqs_py.cpp:193-201: np.array(reg, copy=False) aliases the state."""
    psi = iqs.QubitRegister(3, "base", 0, 0)
    view = np.array(psi, copy=False)
    assert view.shape == (8,) and view[0] == 1
    view[0], view[5] = 0, 1  # host write through the view ...
    assert psi[5] == 1  # ... is what the register sees
    psi.ApplyPauliX(0)  # |101> -> |100>
    assert psi.GetProbability(0) == 0 and psi.GetProbability(2) == 1
    psi[4] = 0.6
    psi[0] = 0.8
    assert abs(psi.ComputeNorm() - 1.0) < 1e-15


def test_apply_1q_gate_fixture(iqs):
    """unit_test/include/apply_1q_gate_test.hpp:48-180 (4 qubits, gates on qubit 3)."""
    n = 4
    # Hadamard on |0000>
    psi = iqs.QubitRegister(n, "base", 0, 0)
    psi.ApplyHadamard(3)
    a = amps(psi)
    assert abs(a[0] - 1 / math.sqrt(2)) < ACC and abs(a[8] - 1 / math.sqrt(2)) < ACC and abs(np.linalg.norm(a) - 1) < ACC
    # RX(0.83) on |0>: cos(t/2)|0> - i sin(t/2)|1>
    t = 0.83
    psi = iqs.QubitRegister(n, "base", 0, 0)
    psi.ApplyRotationX(3, t)
    a = amps(psi)
    assert abs(a[0] - math.cos(t / 2)) < ACC and abs(a[8] - (-1j * math.sin(t / 2))) < ACC
    # RY(0.75) on |1>: -sin(t/2)|0> + cos(t/2)|1>
    t = 0.75
    psi = iqs.QubitRegister(n, "base", 8, 0)
    psi.ApplyRotationY(3, t)
    a = amps(psi)
    assert abs(a[0] + math.sin(t / 2)) < ACC and abs(a[8] - math.cos(t / 2)) < ACC
    # RZ(0.35) on |+>
    t = 0.35
    psi = iqs.QubitRegister(n, "base", 0, 0)
    psi.ApplyHadamard(3)
    psi.ApplyRotationZ(3, t)
    a = amps(psi)
    f = 1 / math.sqrt(2)
    assert abs(a[0] - f * complex(math.cos(t / 2), -math.sin(t / 2))) < ACC and abs(a[8] - f * complex(math.cos(t / 2), math.sin(t / 2))) < ACC
    # custom gate: the fixed G (apply_1q_gate_test.hpp:241-245)
    G = C.G_FIXED.view(np.complex128).reshape(2, 2)
    psi = iqs.QubitRegister(n, "base", 0, 0)
    psi.Apply1QubitGate(3, G)
    a = amps(psi)
    assert abs(a[0] - G[0, 0]) < ACC and abs(a[8] - G[1, 0]) < ACC


def test_single_qubit_gates_fixture(iqs):
    """unit_test/include/single_qubit_gates_test.hpp:43-120: 10 qubits, Pauli strings on basis states."""
    n = 10
    psi = iqs.QubitRegister(n, "base", 0, 0)
    for q in (0, 3, 9):
        psi.ApplyPauliX(q)
    idx = (1 << 0) | (1 << 3) | (1 << 9)
    a = amps(psi)
    assert a[idx] == 1 and np.count_nonzero(a) == 1
    psi.ApplyPauliY(3)  # Y|1> = -i|0>
    a = amps(psi)
    assert a[idx & ~(1 << 3)] == -1j and np.count_nonzero(a) == 1
    psi.ApplyPauliZ(9)  # Z|1> = -|1>
    assert amps(psi)[idx & ~(1 << 3)] == 1j
    for q in range(n):
        assert psi.GetProbability(q) == (1.0 if q in (0, 9) else 0.0)


def test_expectation_values_fixture(iqs):
    """unit_test/include/expectation_values_test.hpp:121-239, tolerance 1e-14."""
    n = 6
    psi = iqs.QubitRegister(n, "base", 0, 0)  # |000000>
    psi.ApplyPauliX(1)  # qubit 1 in |1>
    psi.ApplyHadamard(2)  # qubit 2 in |+>
    psi.ApplyHadamard(3)
    psi.ApplyPauliZ(3)  # qubit 3 in |->
    psi.ApplyHadamard(4)
    psi.ApplyPauliSqrtZ(4)  # qubit 4 in |+i> (eigenstate of Y, +1)
    tol = 1e-14
    assert abs(psi.ExpectationValueZ(0) - 1) < tol and abs(psi.ExpectationValueZ(1) + 1) < tol
    assert abs(psi.ExpectationValueX(2) - 1) < tol and abs(psi.ExpectationValueX(3) + 1) < tol
    assert abs(psi.ExpectationValueY(4) - 1) < tol
    assert abs(psi.ExpectationValueX(0)) < tol and abs(psi.ExpectationValueY(0)) < tol and abs(psi.ExpectationValueZ(2)) < tol
    assert abs(psi.ExpectationValue([0, 1], [3, 3], 1.0) + 1) < tol  # ZZ
    assert abs(psi.ExpectationValue([2, 3], [1, 1], 1.0) + 1) < tol  # XX
    assert abs(psi.ExpectationValue([1, 2, 4], [3, 1, 2], 1.0) + 1) < tol  # ZXY
    assert abs(psi.ExpectationValue([0, 1, 2, 3, 4], [3, 3, 1, 1, 2], 2.0) - 2.0) < tol
    assert abs(psi.ExpectationValue([0, 2], [1, 1], 1.0)) < tol
    assert psi.ExpectationValue([], [], 0.7) == 0.7
    assert abs(psi.ComputeNorm() - 1) < 1e-14  # the state is restored


def test_measure_fixture(iqs):
    """unit_test/include/measure_test.hpp:37-57 and tutorials/get_started_with_IQS.cpp:256-295."""
    n = 6
    psi = iqs.QubitRegister(n, "++++", 0, 0)
    for q in range(n):
        assert abs(psi.GetProbability(q) - 0.5) < 1e-14
    psi.CollapseQubit(2, True)
    psi.Normalize()
    assert abs(psi.GetProbability(2) - 1) < 1e-14 and abs(psi.ComputeNorm() - 1) < 1e-14
    assert psi.IsClassicalBit(2) and psi.GetClassicalValue(2) is True and not psi.IsClassicalBit(3)
    psi.CollapseQubit(3, False)
    psi.Normalize()
    assert psi.GetClassicalValue(3) is False


def test_toffoli_counts_and_truth_table(iqs):
    """ApplyToffoli = 5 two-qubit gates (gate_counter_test.hpp:150-172) and flips the target iff both controls are 1."""
    n = 5
    for c1 in (0, 1):
        for c2 in (0, 1):
            psi = iqs.QubitRegister(n, "base", (c1 << 0) | (c2 << 3), 0)
            psi.ApplyToffoli(0, 3, 4)
            want = (c1 << 0) | (c2 << 3) | ((c1 & c2) << 4)
            a = amps(psi)
            assert abs(a[want] - 1) < 1e-15 and abs(np.linalg.norm(a) - 1) < 1e-15


def test_permute_fixture(iqs):
    """unit_test/include/qureg_permute_test.hpp:60-259: amplitude j -> position given by the new map."""
    n = 6
    psi = iqs.QubitRegister(n, "base", 0, 0)
    for j in range(1 << n):
        psi[j] = complex(j, 0)
    new_map = [4, 1, 2, 5, 0, 3]
    psi.PermuteQubits(new_map, "direct")
    assert psi.GetQubitMap() == new_map
    a = amps(psi)
    for j in range(1 << n):
        data_index = sum(((j >> q) & 1) << new_map[q] for q in range(n))
        assert a[data_index] == j  # bit-exact data movement
        assert psi.GetGlobalAmplitude(j) == j  # program-order view is unchanged
    psi.PermuteQubits(list(range(n)), "direct")
    assert np.array_equal(amps(psi), np.arange(1 << n))


def test_random_state_matches_reference_stream(iqs):
    """Initialize("rand") draws from the same mt19937 stream layout as the reference (qureg_init.cpp:256-332)."""
    import random  # noqa: F401

    rng = iqs.RandomNumberGenerator()
    rng.SetSeedStreamPtrs(777)
    psi = iqs.QubitRegister(10, "base", 0, 0)
    psi.SetRngPtr(rng)
    psi.Initialize("rand", 0)
    a = amps(psi)
    assert abs(np.linalg.norm(a) - 1) < 1e-14
    # pool stream seeded with 777: libstdc++ mt19937 + uniform_real_distribution<double>(0,1), a + (b-a)*u
    # numpy's MT19937 yields the same 32-bit outputs; uniform_real_distribution<double> consumes two per double
    mt = np.random.MT19937()
    mt._legacy_seeding(777)
    raw = mt.random_raw(4 << 10).astype(np.float64)
    u = (raw[0::2] + raw[1::2] * 4294967296.0) / 18446744073709551616.0
    v = -1.0 + 2.0 * u
    ref = v[0::2] + 1j * v[1::2]
    ref /= np.linalg.norm(ref)
    assert np.max(np.abs(a - ref)) < 1e-14


def test_chi_matrix_binding_and_channels(iqs):
    """CM4x4 / CM16x16, SolveEigenSystem and ApplyChannel from Python (reference intelqs_py.cpp:94-163,
    263-271; usage as in notebooks/chi_matrix_with_mpi.py:24-49)."""
    p = 0.1
    # amplitude-damping chi of the reference notebook (has a negative eigenvalue)
    chi = iqs.CM4x4()
    chi[0, 0] = (1 + np.sqrt(1 - p)) ** 2
    chi[0, 3] = p
    chi[3, 0] = p
    chi[3, 3] = (1 - np.sqrt(1 - p)) ** 2
    chi[1, 1] = p ** 2
    chi[1, 2] = -1j * p ** 2
    chi[2, 1] = +1j * p ** 2
    chi[2, 2] = p ** 2
    assert chi[1, 2] == -1j * p ** 2
    chi.SolveEigenSystem()
    E = np.array(chi.GetEigenValues()).real
    W = np.array(chi.GetEigenVectors())  # W[k, i]
    probs = np.array(chi.GetEigenProbabilities())
    total = np.abs(E).sum()
    assert np.all(np.diff(E) >= 0) and abs(probs.sum() - 1) < 1e-14
    want = np.array([[chi[i, j] for j in range(4)] for i in range(4)])
    got = sum(E[k] / total * np.outer(W[k], W[k].conj()) for k in range(4))
    assert np.max(np.abs(got - want)) < 1e-13
    with pytest.raises(IndexError):
        chi[4, 0]

    # dephasing rho' = (1-p) rho + p Z rho Z on |+>: every trajectory stays |+> or |->, <X> decays as (1-2p)^t
    chi = iqs.CM4x4()
    chi[0, 0] = 1 - p
    chi[3, 3] = p
    chi.SolveEigenSystem()
    rng = iqs.RandomNumberGenerator()
    rng.SetSeedStreamPtrs(7777)
    plus = iqs.QubitRegister(3, "base", 0, 0)
    plus.ApplyHadamard(1)
    steps, ensemble, mean = 5, 300, 0.0
    for _ in range(ensemble):
        psi = iqs.QubitRegister(plus)
        psi.SetRngPtr(rng)
        for _t in range(steps):
            psi.ApplyChannel(1, chi)
        ov = abs(psi.ComputeOverlap(plus)) ** 2
        assert min(abs(ov), abs(ov - 1)) < 1e-12 and abs(psi.ComputeNorm() - 1) < 1e-12
        mean += (2 * ov - 1) / ensemble
        assert psi.GetOverallSignOfChannels() == 1
    assert abs(mean - (1 - 2 * p) ** steps) < 0.2

    # two-qubit channel: chi of the ideal CZ = 1/2 (II + IZ + ZI - ZZ) reproduces the gate
    chi2 = iqs.CM16x16()
    idx, v = [0, 3, 12, 15], [0.5, 0.5, 0.5, -0.5]
    for a in range(4):
        for b in range(4):
            chi2[idx[a], idx[b]] = v[a] * v[b]
    chi2.SolveEigenSystem()
    psi = iqs.QubitRegister(5, "++++", 0, 0)
    ideal = iqs.QubitRegister(5, "++++", 0, 0)
    psi.SetRngPtr(rng)
    psi.ApplyChannel(0, 3, chi2)
    ideal.ApplyCPauliZ(0, 3)
    assert abs(abs(ideal.ComputeOverlap(psi)) ** 2 - 1) < 1e-12


def test_reference_binding_source_compiled_unchanged():
    """The f1 drop-in proof: the reference's pybind11/intelqs_py.cpp, compiled UNCHANGED against
    intel-qs_b200/include + libiqs.so (oracle/Makefile, target _ref/dropin/pybind), runs the reference's
    unit_test/import_iqs.py scenario and the README circuit; the amplitudes equal the ones of the
    engine's own module.  (Separate process: both modules are called intelqs_py.)"""
    import glob
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    moddir = os.path.join(root, "oracle", "_ref", "dropin", "pybind")
    assert glob.glob(os.path.join(moddir, "intelqs_py*.so")), "oracle/_ref/dropin/pybind is missing: run __graft_entry__.build() where /root/reference exists"
    script = r"""
import sys
sys.path.insert(0, sys.argv[1])
import numpy as np
import intelqs_py as iqs
assert iqs.__file__.startswith(sys.argv[1])
iqs.EnvInit()
rank = iqs.MPIEnvironment.GetRank()
psi = iqs.QubitRegister(2, "base", 0, 0)   # unit_test/import_iqs.py
print("The IQS library was successfully imported and initialized.")
n = 6
psi = iqs.QubitRegister(n, "base", 1, 0)
for q in range(n):
    psi.ApplyHadamard(q)
psi.ApplyCPauliX(0, 3)
psi.ApplyRotationY(2, 0.37)
psi.ApplyToffoli(1, 2, 4)
G = np.zeros((2, 2), dtype=np.complex128)
G[0, 0], G[0, 1], G[1, 0], G[1, 1] = 0.5922 + 0.4596j, -0.0387 - 0.6608j, -0.1305 + 0.6490j, 0.4966 + 0.5621j
psi.Apply1QubitGate(5, G)
p = psi.GetProbability(4)
amps = np.array([psi[i] for i in range(1 << n)])
np.save(sys.argv[2], np.concatenate([amps, [p]]))
iqs.EnvFinalize()
"""
    import tempfile

    outs = []
    with tempfile.TemporaryDirectory() as td:
        for k, d in enumerate((moddir, os.path.join(root, "intel-qs_b200", "lib"))):
            out = os.path.join(td, f"o{k}.npy")
            r = subprocess.run([sys.executable, "-c", script, d, out], capture_output=True, text=True, timeout=300)
            assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
            assert "successfully imported" in r.stdout
            outs.append(np.load(out))
    assert np.array_equal(outs[0], outs[1])
    assert abs(np.sum(np.abs(outs[0][:-1]) ** 2) - 1.0) < 0.05  # (the 4-digit matrix above is unitary to 1e-2 only)
