"""Capture the stdout of the reference's own programs built against the UNMODIFIED reference library
(oracle/_ref/refbin/*, see oracle/Makefile) as fixtures for tests/test_examples_dropin_gpu.py.
Run in the build container: python tests/golden/make_example_outputs.py"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
BIN = os.path.join(ROOT, "oracle", "_ref", "refbin")

COMMANDS = {
    "grover_4qubit": [],
    "expect_value_test": [],
    "heisenberg_dynamics": ["8"],
    "quantum_fourier_transform": ["8"],
    "test_of_custom_gates": ["10"],
    "benchgates": ["10"],
    "specv2_bench": ["10", "1"],
    "get_started_with_IQS": [],
    "communication_reduction_via_qubit_reordering": ["22"],
    "circuit_with_noise_gates": ["8"],
    "get_started_with_noisy_IQS": [],
    # (noise_via_chi_matrix needs the reference configured with Eigen: no fixture, see the test)
}

# the QASM interpreter reads its program from stdin (interface/src/interface_api_qasm.cpp:109-125)
QASM = ".malloc 3\nH q0\nCNOT q0,q1\nT q1\nS q2\nX q2\nTdag q0\nMeasZ q0\nMeasZ q1\nMeasZ q2\n.version\n.free\n\n"

if __name__ == "__main__":
    env = dict(os.environ, OMP_NUM_THREADS="4")
    r = subprocess.run([os.path.join(BIN, "iqs_interface")], input=QASM, capture_output=True, text=True, env=env, timeout=600)
    open(os.path.join(HERE, "examples", "iqs_interface.txt"), "w").write(f"EXIT {r.returncode}\n" + r.stdout)
    print("iqs_interface exit", r.returncode)
    r = subprocess.run([os.path.join(BIN, "qaoa_check")], capture_output=True, text=True, env=env, timeout=600)
    open(os.path.join(HERE, "examples", "qaoa_check.txt"), "w").write(f"EXIT {r.returncode}\n" + r.stdout)
    print("qaoa_check exit", r.returncode)
    r = subprocess.run([os.path.join(BIN, "noise_check")], capture_output=True, text=True, env=env, timeout=600)
    open(os.path.join(HERE, "examples", "noise_check.txt"), "w").write(f"EXIT {r.returncode}\n" + r.stdout)
    print("noise_check exit", r.returncode)
    for name, args in COMMANDS.items():
        r = subprocess.run([os.path.join(BIN, name)] + args, capture_output=True, text=True, env=env, timeout=600)
        # (benchgates ends with `return 1` also on success; the exit code is part of the fixture)
        open(os.path.join(HERE, "examples", name + ".txt"), "w").write(f"EXIT {r.returncode}\n" + r.stdout)
        print(name, "exit", r.returncode, len(r.stdout.splitlines()), "lines")
