"""The reference's OWN programs (examples/, tutorials/, benchmarks/), compiled without any source
change against intel-qs_b200/include + libiqs.so (oracle/Makefile, target _ref/dropin/bin/*), run on
the GPU and must print what they print with the reference library (fixtures tests/golden/examples/*.txt,
captured by tests/golden/make_example_outputs.py from the reference build).

Lines that carry wall-clock times or build-specific notices are ignored; numbers are compared with
the tolerance the program's own checks use."""
import os
import re
import subprocess

import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
BIN = os.path.join(ROOT, "oracle", "_ref", "dropin", "bin")
GOLD = os.path.join(HERE, "golden", "examples")

NUM = re.compile(r"[-+]?(?:\d+\.\d*|\.\d+|\d+)(?:[eE][-+]?\d+)?")
SKIP = re.compile(r"MPI not enabled|seconds|Simulation time|OMP number of threads|statistics|IqsMPI|INTELQS_HAS_MPI|Fusion is|Compiler flags|-->|Time |time ", re.I)


def run(name, args, stdin=None):
    exe = os.path.join(BIN, name)
    if not os.path.exists(exe):
        pytest.skip(f"{exe} not built (needs /root/reference at build time)")
    r = subprocess.run([exe] + args, input=stdin, capture_output=True, text=True, timeout=600)
    return r.returncode, r.stdout, r.stderr


def golden(name):
    txt = open(os.path.join(GOLD, name + ".txt")).read().splitlines()
    assert txt[0].startswith("EXIT ")
    return int(txt[0].split()[1]), txt[1:]


def keep(lines):
    return [l.rstrip() for l in lines if l.strip() and not SKIP.search(l)]


def same(a, b, tol):
    """same text skeleton, numbers within tol"""
    if NUM.sub("#", a).split() != NUM.sub("#", b).split():
        return False
    xa, xb = [float(x) for x in NUM.findall(a)], [float(x) for x in NUM.findall(b)]
    return len(xa) == len(xb) and all(abs(p - q) <= tol for p, q in zip(xa, xb))


def compare(name, args, tol=1e-7, rc_must_match=True, stdin=None, until=None):
    """until: a marker line; what the drop-in build prints from there on has no reference counterpart"""
    want_rc, want = golden(name)
    rc, out, err = run(name, args, stdin)
    if rc_must_match:
        assert rc == want_rc, f"{name}: exit {rc}, reference exits {want_rc}\n{err[-1500:]}"
    lines = out.splitlines()
    if until is not None:
        cut = [i for i, l in enumerate(lines) if l.startswith(until)]
        assert cut, f"{name}: marker {until!r} not printed"
        lines, extra = lines[: cut[0]], lines[cut[0] + 1 :]
    got, want = keep(lines), keep(want)
    assert len(got) == len(want), f"{name}: {len(got)} lines vs {len(want)} in the reference output\n" + "\n".join(got[-15:])
    for g, w in zip(got, want):
        assert same(g, w, tol), f"{name}:\n  ours: {g}\n  ref : {w}"
    return extra if until is not None else None


def test_grover_4qubit():
    compare("grover_4qubit", [])  # final amplitudes -0.1875 x15, -0.6875 at |0100>


def test_expect_value_test():
    compare("expect_value_test", [])  # 1, 0, -1, -1


def test_heisenberg_dynamics_8():
    compare("heisenberg_dynamics", ["8"], tol=1e-11)


def test_test_of_custom_gates():
    compare("test_of_custom_gates", ["10"])


def test_get_started_tutorial():
    compare("get_started_with_IQS", [])  # seeded measurement outcome included


def test_benchgates_self_check():
    """asserts overlap - 1 < 1e-13 between generic and specialised kernels after every gate (benchgates.cpp:140,157)."""
    compare("benchgates", ["10"])


def test_specv2_bench_self_check():
    rc, out, err = run("specv2_bench", ["10", "1"])
    assert rc == 0, err[-1500:]
    assert "State comparison test passed for spec v1 & v2" in out


def test_communication_reduction_example():
    rc, out, err = run("communication_reduction_via_qubit_reordering", ["22"])
    assert rc == 0, err[-1500:]
    m = re.search(r"Squared overlap of states at the end of the two simulations = ([-+0-9.eE]+)", out)
    assert m and abs(float(m.group(1)) - 1.0) < 1e-12


def test_quantum_fourier_transform_example():
    """BASELINE configs[0] program itself (at 8 qubits; 20 segfaults in the reference, SURVEY.md top table):
    QFT of a seeded random state vs the example's classical DFT, in single and double precision."""
    rc, out, err = run("quantum_fourier_transform", ["8"])
    assert rc == 0, err[-1500:]
    sp = re.search(r"SP::qufft error vs classical max\(absdiff: ([0-9.eE+-]+)", out)
    dp = re.search(r"DP::qufft error vs classical max\(absdiff: ([0-9.eE+-]+)", out)
    assert sp and float(sp.group(1)) < 1e-5
    assert dp and float(dp.group(1)) < 1e-13
    # same seeded input, same arithmetic: the DP error equals the reference's to the printed digits
    _, want = golden("quantum_fourier_transform")
    wdp = [l for l in want if l.startswith("DP::qufft")][0]
    assert abs(float(dp.group(1)) - float(re.search(r"absdiff: ([0-9.eE+-]+)", wdp).group(1))) < 1e-15


def test_qasm_interface():
    """interface/src/*.cpp, the stdin QASM interpreter, relinked unchanged."""
    qasm = ".malloc 3\nH q0\nCNOT q0,q1\nT q1\nS q2\nX q2\nTdag q0\nMeasZ q0\nMeasZ q1\nMeasZ q2\n.version\n.free\n\n"
    compare("iqs_interface", [], stdin=qasm)


def test_qaoa_features():
    """iqs::qaoa::* (SURVEY.md 8f row 3): MaxCut cost vectors (integer and weighted, permuted qubit
    order), QAOA layers, cost expectation values and the three histograms, against the reference's
    own qaoa_features.cpp run on the CPU."""
    compare("qaoa_check", [], tol=2e-12)


def test_circuit_with_noise_gates_example():
    """examples/circuit_with_noise_gates.cpp: 2 x 200 stochastic circuits through iqs::NoisyQureg
    (include/NoisyQureg.hpp); the seeded std::default_random_engine makes the run reproducible, so
    the averaged overlaps must print as in the reference build."""
    compare("circuit_with_noise_gates", ["8"], tol=2e-6)


def test_noisy_tutorial():
    """tutorials/get_started_with_noisy_IQS.cpp: ApplyNoiseGate on the RNG stream "state", 2 x 200
    trajectories, pool sums (the program ends with `return 1`)."""
    compare("get_started_with_noisy_IQS", [], tol=2e-6)


def test_noise_scenario():
    """tests/noise_check.cpp: every NoisyQureg method, ApplyNoiseGate, ApplyChannel with the closed-form
    Hadamard eigensystem -- amplitudes against the reference build; then (B200 build only, the
    reference needs Eigen) the depolarising channel and a two-qubit channel through the eigen-solver."""
    extra = compare("noise_check", [], tol=2e-12, until="==== eigen-solver part")
    assert any(l.startswith("depolarising OK") for l in extra), extra
    assert any(l.startswith("CZ channel") and l.endswith("OK") for l in extra), extra
    assert not any("FAILED" in l for l in extra)


def test_noise_via_chi_matrix_example():
    """examples/noise_via_chi_matrix.cpp: the reference build aborts here (SolveEigenSystem asserts
    without Eigen), so the program's own statement is the check: the channel is an ideal Hadamard,
    the noisy ensemble must equal the noiseless circuit."""
    rc, out, err = run("noise_via_chi_matrix", ["8"])
    assert rc == 1, err[-1500:]  # the program ends with `return 1`
    ov = re.search(r"Overlap-squared between ideal and noisy states = ([-+0-9.eE]+)", out)
    p0 = re.search(r"in the noiseless case = ([-+0-9.eE]+)", out)
    p1 = re.search(r"with noise = ([-+0-9.eE]+)", out)
    assert ov and abs(float(ov.group(1)) - 1.0) < 1e-6
    assert p0 and p1 and abs(float(p0.group(1)) - float(p1.group(1))) < 1e-6
