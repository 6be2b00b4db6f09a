"""Distributed parity on 2/4/8 GPUs of one box (skipped when fewer are visible): the drop-in driver
is launched with tools/iqsrun (one process per GPU, NCCL bootstrap, cudaIpc peer memory); the global
state gathered from the shards must equal the single-rank oracle."""
import os
import subprocess
import sys

import numpy as np
import pytest

from pkg import circuits as C
from progs import random_program

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "intel-qs_b200", "bin", "iqs_b200_driver")
IQSRUN = os.path.join(ROOT, "tools", "iqsrun")
TOL = 1e-12


def gpu_count():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True).stdout
        return sum(1 for l in out.splitlines() if l.startswith("GPU "))
    except Exception:
        return 0


NGPU = gpu_count()


def run_ranks(oracle, nranks, prog, state=None, **kw):
    launcher = [sys.executable, IQSRUN, "-n", str(nranks), "--timeout", "300"]
    return oracle.run_driver(DRIVER, prog, state=state, launcher=launcher, **kw)


def random_unitary_for(rng):
    from progs import random_unitary

    return random_unitary(rng)


def need(n):
    if NGPU < n:
        pytest.skip(f"needs {n} GPUs, {NGPU} visible")


@pytest.mark.parametrize("nranks", [2, 4, 8])
@pytest.mark.parametrize("mode", ["placement", "eager", "pairkernels"])
def test_gates_and_permutations_sharded_bit_exact(oracle, nranks, mode):
    """Every gate kind on local and global qubits, interleaved with local / global / mixed qubit
    permutations (qureg_permute_test.hpp): the gathered state equals the oracle bit for bit.
    mode: the placement layer with its look-ahead queue (default), without look-ahead, or switched
    off (gates on rank bits run as peer-memory pair kernels, the reference's HP_Distrpair replaced
    one to one)."""
    need(nranks)
    env = {"placement": {}, "eager": {"IQS_B200_LOOKAHEAD": "0"}, "pairkernels": {"IQS_B200_PLACEMENT": "0"}}[mode]
    n, seed = 12, 2
    rng = np.random.default_rng(nranks)
    prog = random_program(n, 200, seed, toffoli=True)
    prog.permute(list(rng.permutation(n)))
    prog.extend(random_program(n, 80, seed + 1))
    prog.permute(list(range(n))[::-1])  # reverses the global qubits too: in-place shard exchange
    prog.named1(C.H, 0).named2(C.CX, 0, n - 1)
    prog.permute(list(rng.permutation(n)))
    prog.extend(random_program(n, 40, seed + 2))
    psi = C.random_state(n, seed)
    want, _, wmap = oracle.run_program(n, psi, prog.ops)
    got = run_ranks(oracle, nranks, prog, state=psi, extra_env=env)
    assert np.array_equal(got["map"], wmap)
    err = np.max(np.abs(got["state"] - want))
    assert err <= TOL, f"{nranks} ranks: max |amp - oracle| = {err}"
    assert np.array_equal(got["state"], want)


@pytest.mark.parametrize("nranks", [2, 4, 8])
def test_scalars_and_measurement_sharded(oracle, nranks):
    need(nranks)
    n, seed = 10, 4
    prog = random_program(n, 120, seed)
    for q in range(n):
        prog.prob(q)
    prog.expect([0, n - 1], [1, 3]).expect([n - 1, n - 2, 1], [2, 1, 3]).expect1(n - 1, 1).expect1(n - 1, 2).expect1(0, 3).norm()
    prog.collapse(n - 1, 1).normalize().norm().collapse(0, 0).normalize().prob(n - 1)
    psi = C.random_state(n, seed)
    want, wsc, _ = oracle.run_program(n, psi, prog.ops)
    got = run_ranks(oracle, nranks, prog, state=psi)
    assert np.max(np.abs(got["scalars"] - wsc)) <= TOL
    assert np.max(np.abs(got["state"] - want)) <= TOL


@pytest.mark.parametrize("nranks", [2, 8])
def test_qft_sharded(oracle, nranks):
    """BASELINE configs[2] at a size the oracle finishes in seconds."""
    need(nranks)
    n = 16
    prog = C.qft(n)
    psi = C.random_state(n, seed=777)
    want, _, _ = oracle.run_program(n, psi, prog.ops)
    got = run_ranks(oracle, nranks, prog, state=psi)
    assert np.max(np.abs(got["state"] - want)) <= TOL


@pytest.mark.parametrize("nranks,placement", [(2, "1"), (2, "0"), (8, "1")])
def test_fusion_sharded(oracle, nranks, placement):
    need(nranks)
    n = 13
    prog = C.Program(n).mode(C.FUSION_ON, 8)
    prog.extend(random_program(n, 200, 5, kinds="basic"))
    prog.mode(C.FUSION_OFF)
    psi = C.random_state(n, 6)
    want, _, _ = oracle.run_program(n, psi, prog.ops)
    got = run_ranks(oracle, nranks, prog, state=psi, extra_env={"IQS_B200_PLACEMENT": placement})
    assert np.array_equal(got["state"], want)


@pytest.mark.parametrize("nranks", [2, 4, 8])
def test_placement_layer_reads_between_moves(oracle, nranks):
    """Qubits travel between local bits and rank bits while the program keeps reading: probabilities,
    expectation values, single amplitudes, collapse, SWAPs with global qubits (relabelled, no data
    moved), a permutation in the middle (forces the reference's layout back) -- scalars to 1e-12,
    final gathered state bit for bit."""
    need(nranks)
    n, seed = 12, 21
    rng = np.random.Generator(np.random.MT19937(seed))
    prog = C.Program(n)
    for rnd in range(6):
        for q in range(n - 1, -1, -1):
            prog.gate1(q, random_unitary_for(rng))
        prog.named2(C.SWAP, 0, n - 1).named2(C.SWAP, n - 2, n - 1).named2(C.ISWAP, 1, n - 1).named2(C.SQRTISWAP, n - 1, n - 3)
        prog.named2(C.CX, n - 1, 2).named2(C.CX, 2, n - 1).named2(C.CPHASE, n - 1, n - 2, 0.3).named1(C.T, n - 1)
        prog.prob(n - 1).prob(0).expect([n - 1, 0], [1, 2]).expect1(n - 2, 2).get_amp(int(rng.integers(0, 1 << n))).get_amp((1 << n) - 1)
        if rnd == 2:
            prog.permute([int(x) for x in rng.permutation(n)])
        if rnd == 4:
            prog.named1(C.H, n - 1).collapse(n - 1, 1).normalize().emuswap(3, n - 1)
        prog.toffoli(n - 1, 4, n - 2)
    psi = C.random_state(n, seed)
    want, wsc, wmap = oracle.run_program(n, psi, prog.ops)
    # IQS_B200_ONE_SWEEP=0: expectation values rotate the state into the observable's basis and back as
    # the reference does (two roundings per amplitude), so the final state can be compared bit for bit
    got = run_ranks(oracle, nranks, prog, state=psi, extra_env={"IQS_B200_ONE_SWEEP": "0"})
    assert np.array_equal(got["map"], wmap)
    assert got["scalars"].size == wsc.size and np.max(np.abs(got["scalars"] - wsc)) <= TOL
    assert np.array_equal(got["state"], want), np.max(np.abs(got["state"] - want))
    # default: read-only expectation values and cached marginals -- the state is not touched by them
    got = run_ranks(oracle, nranks, prog, state=psi)
    assert np.array_equal(got["map"], wmap)
    assert got["scalars"].size == wsc.size and np.max(np.abs(got["scalars"] - wsc)) <= TOL
    assert np.max(np.abs(got["state"] - want)) <= TOL


@pytest.mark.parametrize("nranks", [2, 4])
def test_two_qubit_gate_on_global_qubits(oracle, nranks):
    """Apply2QubitGate with one or both qubits global (the reference asserts a single rank,
    qureg_apply2qubitgate.cpp:23): global positions are exchanged with local ones by exact moves,
    so the gathered state equals the single-rank oracle bit for bit."""
    need(nranks)
    n = 10
    rng = np.random.Generator(np.random.MT19937(17))
    prog = C.Program(n)
    pairs = [(n - 1, 0), (0, n - 1), (n - 1, n - 2), (n - 2, n - 1), (3, 5), (n - 1, n - 3), (n - 4, n - 1)]
    pairs += [tuple(int(x) for x in rng.permutation(n)[:2]) for _ in range(12)]
    for qh, ql in pairs:
        a = rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4))
        u, _ = np.linalg.qr(a)
        prog.gate2(qh, ql, u)
        prog.named1(C.H, int(rng.integers(0, n)))
    psi = C.random_state(n, 8)
    want, _, _ = oracle.run_program(n, psi, prog.ops)
    got = run_ranks(oracle, nranks, prog, state=psi)
    assert np.array_equal(got["state"], want)


@pytest.mark.parametrize("nranks", [2, 4])
def test_pool_of_states_one_state_per_gpu(nranks):
    """iqs::mpi::Environment::UpdateStateComm(num_states = number of GPUs): an ensemble of noisy
    trajectories, one single-GPU state per rank, IncoherentSumOverAllStatesOfPool across them
    (reference noisy_simulation_test.hpp:52-158) -- tests/pool_check.cpp under tools/iqsrun."""
    need(nranks)
    exe = os.path.join(ROOT, "intel-qs_b200", "bin", "pool_check")
    assert os.path.exists(exe), "intel-qs_b200/bin/pool_check is missing: run __graft_entry__.build()"
    r = subprocess.run([sys.executable, IQSRUN, "-n", str(nranks), "--timeout", "300", exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "ALL OK" in r.stdout and "FAILED" not in r.stdout
    assert "OK one_state_at_a_time" in r.stdout and "OK one_state_per_rank" in r.stdout


def _free_gib_per_gpu():
    try:
        out = subprocess.run(["nvidia-smi", "--query-gpu=memory.free", "--format=csv,noheader,nounits"], capture_output=True, text=True).stdout
        return min(float(x) for x in out.split()) / 1024.0
    except Exception:
        return 0.0


@pytest.mark.parametrize("nranks,n", [(2, 20), (4, 21), (8, 22), (2, 33), (4, 34), (8, 35)])
@pytest.mark.parametrize("fused", [False, True])
def test_closed_form_amplitudes_sharded(oracle, nranks, n, fused):
    """An independent answer at sizes no CPU reference holds (35 qubits on 8 GPUs = BASELINE configs[3]
    size): tests/closed_form.py -- product state, CNOT chain, controlled phases -- with sampled
    GetGlobalAmplitude reads that force every global bit, against the analytic formula at 1e-12."""
    need(nranks)
    M = n - int(np.log2(nranks))
    if M >= 30 and _free_gib_per_gpu() < (16 * (1 << M) >> 30) + 4:
        pytest.skip(f"needs {(16 * (1 << M) >> 30) + 4} GiB free per GPU")
    from closed_form import ClosedForm

    cf = ClosedForm(n, seed=100 + n)
    samples = cf.samples(10000 if M >= 30 else 2000)
    got = run_ranks(oracle, nranks, cf.program(C, samples, fused=fused), init=1, base_index=0, want_state=False)
    sc = got["scalars"]
    assert sc.size == 2 * len(samples)
    err, scale = cf.check(samples, sc[0::2] + 1j * sc[1::2], TOL)
    print(f"closed form {n} qubits on {nranks} GPUs fused={fused}: {len(samples)} amplitudes, max |d| = {err:.3e} (largest |amp| {scale:.3e}), {got['seconds']:.2f} s")


def test_barrier_gives_up_instead_of_hanging():
    """One rank never reaches the rendezvous: the other gets IQSB_ERR_PEER after the deadline
    (IQS_B200_BARRIER_TIMEOUT_S) -- the kernel does not spin for ever."""
    need(2)
    import subprocess

    script = os.path.join(ROOT, "tests", "barrier_timeout_check.py")
    r = subprocess.run([sys.executable, IQSRUN, "-n", "2", "--timeout", "120", sys.executable, script], capture_output=True, text=True, timeout=180)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "TIMEOUT_OK" in r.stdout
