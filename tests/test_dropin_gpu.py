"""The drop-in proof: oracle/driver.cpp -- written against the reference's public C++ API and
compiled UNCHANGED against intel-qs_b200/include + libiqs.so -- replays programs on the GPU; results
are compared with the CPU oracle (which is pinned bit-exact to the compiled reference).

Covers the host-side dispatch of iqs::QubitRegister (qubit->position maps, named-gate matrices,
diagonal shortcuts, fusion queue, permutations, expectation values, measurement)."""
import os

import numpy as np
import pytest

from pkg import circuits as C
from progs import random_program

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "intel-qs_b200", "bin", "iqs_b200_driver")
TOL = 1e-12


def run_gpu(oracle, prog, state=None, **kw):
    assert os.path.exists(DRIVER), "intel-qs_b200/bin/iqs_b200_driver is missing: run __graft_entry__.build()"
    return oracle.run_driver(DRIVER, prog, state=state, **kw)


@pytest.mark.parametrize("n,nops,seed", [(3, 150, 11), (5, 300, 12), (10, 400, 13), (14, 300, 14)])
def test_random_programs_match_oracle(oracle, n, nops, seed):
    prog = random_program(n, nops, seed)
    psi = C.random_state(n, seed)
    want, _, wmap = oracle.run_program(n, psi, prog.ops)
    got = run_gpu(oracle, prog, state=psi)
    assert np.array_equal(got["map"], wmap)
    err = np.max(np.abs(got["state"] - want))
    assert err <= TOL, f"max |amp - oracle| = {err}"
    # same operation order as the reference, no FMA: expected to be exactly equal
    assert np.array_equal(got["state"], want)


def test_permuted_register_and_scalars(oracle):
    n, seed = 12, 5
    rng = np.random.default_rng(seed)
    prog = random_program(n, 100, seed)
    prog.permute(list(rng.permutation(n)))
    prog.extend(random_program(n, 100, seed + 1))
    prog.emuswap(2, 7)
    prog.extend(random_program(n, 50, seed + 2))
    for q in range(n):
        prog.prob(q)
    prog.expect([0, 1, 2], [1, 2, 3]).expect([3, 9], [3, 3]).expect([4, 5, 6, 7, 8, 11], [1, 1, 2, 2, 3, 3])
    prog.expect1(1, 1).expect1(2, 2).expect1(3, 3).norm()
    prog.collapse(1, 1).normalize().norm().prob(1)
    psi = C.random_state(n, seed)
    want, wsc, wmap = oracle.run_program(n, psi, prog.ops)
    got = run_gpu(oracle, prog, state=psi)
    assert np.array_equal(got["map"], wmap)
    assert np.max(np.abs(got["scalars"] - wsc)) <= TOL
    assert np.max(np.abs(got["state"] - want)) <= TOL


def test_fusion_matches_unfused_reference_semantics(oracle):
    n = 14
    for log2llc in (4, 11, 13):
        prog = C.Program(n).mode(C.FUSION_ON, log2llc)
        prog.extend(random_program(n, 300, 99 + log2llc, kinds="basic"))
        prog.mode(C.FUSION_OFF)
        psi = C.random_state(n, 3)
        want, _, _ = oracle.run_program(n, psi, prog.ops)
        got = run_gpu(oracle, prog, state=psi)
        assert np.array_equal(got["state"], want), f"log2llc={log2llc}: {np.max(np.abs(got['state'] - want))}"


def test_auto_fusion_env_is_transparent(oracle):
    """IQS_B200_AUTO_FUSION=1: an unchanged program (every op kind, permutations, reductions,
    collapse) runs with registers that fuse from the start; states, scalars and maps are the same."""
    n, seed = 13, 41
    rng = np.random.default_rng(seed)
    prog = random_program(n, 250, seed)
    prog.permute(list(rng.permutation(n)))
    prog.extend(random_program(n, 150, seed + 1))
    for q in (0, 5, n - 1):
        prog.prob(q)
    prog.expect([0, 1, 2], [1, 2, 3]).norm()
    prog.collapse(2, 0).norm()  # (no Normalize: its scalar comes from a reduction whose summation order differs)
    prog.extend(random_program(n, 60, seed + 2))
    psi = C.random_state(n, seed)
    want, wsc, wmap = oracle.run_program(n, psi, prog.ops)
    # the reference's ExpectationValue rotates the state into the observable's basis and back (two
    # roundings per amplitude); IQS_B200_ONE_SWEEP=0 does the same sweeps, so the states are identical
    got = run_gpu(oracle, prog, state=psi, extra_env={"IQS_B200_AUTO_FUSION": "1", "IQS_B200_ONE_SWEEP": "0"})
    assert np.array_equal(got["map"], wmap)
    assert np.max(np.abs(got["scalars"] - wsc)) <= TOL
    assert np.array_equal(got["state"], want)
    # default: read-only expectation values (the state is not touched at all), same numbers to 1e-12
    got = run_gpu(oracle, prog, state=psi, extra_env={"IQS_B200_AUTO_FUSION": "1"})
    assert np.max(np.abs(got["scalars"] - wsc)) <= TOL
    assert np.max(np.abs(got["state"] - want)) <= TOL


def test_one_sweep_reductions_through_the_api(oracle):
    """GetProbability of every qubit (all marginals from one read after the first call), every
    1- and 2-qubit Pauli expectation wrapper and longer strings -- the reference's table
    (unit_test/include/expectation_values_test.hpp:121-239) on a random state, permuted register."""
    n, seed = 11, 17
    rng = np.random.default_rng(seed)
    prog = random_program(n, 120, seed)
    prog.permute(list(rng.permutation(n)))
    for rep in range(2):
        for q in range(n):
            prog.prob(q)
        for o in (1, 2, 3):
            for q in (0, 4, n - 1):
                prog.expect1(q, o)
        for o1 in (1, 2, 3):
            for o2 in (1, 2, 3):
                prog.expect([2, 7], [o1, o2]).expect([n - 1, 0], [o1, o2])
        prog.expect(list(range(n)), [1 + (q % 3) for q in range(n)])
        prog.expect([1, 3, 5, 8], [2, 2, 2, 1])
        for q in range(n):
            prog.prob(n - 1 - q)
        prog.named1(C.H, 3).prob(3).prob(0)  # a gate invalidates the cached marginals
    prog.norm()
    psi = C.random_state(n, seed)
    want, wsc, wmap = oracle.run_program(n, psi, prog.ops)
    got = run_gpu(oracle, prog, state=psi)
    assert len(got["scalars"]) == len(wsc)
    assert np.max(np.abs(got["scalars"] - wsc)) <= TOL
    assert np.max(np.abs(got["state"] - want)) <= TOL
    ref_like = run_gpu(oracle, prog, state=psi, extra_env={"IQS_B200_ONE_SWEEP": "0"})
    assert np.max(np.abs(ref_like["scalars"] - wsc)) <= TOL
    assert np.array_equal(ref_like["state"], want)


def test_spec_modes_do_not_change_results(oracle):
    n = 10
    base = random_program(n, 200, 7)
    psi = C.random_state(n, 1)
    want, _, _ = oracle.run_program(n, psi, base.ops)
    for on in (C.SPEC_ON, C.SPEC2_ON):
        prog = C.Program(n).mode(on).extend(base)
        got = run_gpu(oracle, prog, state=psi)
        assert np.array_equal(got["state"], want)


def test_qft_config1(oracle):
    """BASELINE configs[0]: QFT at 20 qubits from a seeded random state, vs the oracle and vs numpy's FFT."""
    n = 20
    prog = C.qft(n)
    psi = C.random_state(n, seed=777)
    want, _, _ = oracle.run_program(n, psi, prog.ops)
    got = run_gpu(oracle, prog, state=psi)
    assert np.max(np.abs(got["state"] - want)) <= TOL
    # the example checks against a backward DFT scaled by 1/sqrt(N) (quantum_fourier_transform.cpp:143-154)
    fft = np.fft.ifft(psi) * np.sqrt(float(1 << n))
    assert np.max(np.abs(got["state"] - fft)) < 1e-10


def test_heisenberg_golden_8_qubits(oracle):
    """examples/heisenberg_dynamics.cpp at 8 qubits: <Z> before and after one Trotter step
    (golden vector 11, captured from the reference binary)."""
    n = 8
    prog = C.heisenberg_step(n)
    got = run_gpu(oracle, prog, init=1, base_index=1)
    before, after = got["scalars"][:n], got["scalars"][n:]
    assert np.allclose(before[0], -0.866025624917, atol=1e-11)
    assert np.allclose(before[1:], 0.866025624917, atol=1e-11)
    want_after = [-0.861699104030, 0.857404954316, 0.865971854486, 0.866014546362, 0.866014759111, 0.866014760171, 0.866014760176, 0.866014760176]
    assert np.allclose(after, want_after, atol=1e-11)


def test_base_and_plus_initialisation(oracle):
    n = 9
    got = run_gpu(oracle, C.Program(n), init=1, base_index=37)
    want = np.zeros(1 << n, dtype=np.complex128)
    want[37] = 1
    assert np.array_equal(got["state"], want)
    got = run_gpu(oracle, C.Program(n), init=2)
    assert np.array_equal(got["state"], np.full(1 << n, 1 / np.sqrt(float(1 << n)), dtype=np.complex128))


def test_device_memory_mode_with_host_mirror(oracle):
    """IQS_B200_MEM=device: operator[] goes through the chunked host mirror (multi-rank code path)."""
    n = 13
    prog = random_program(n, 120, 21)
    psi = C.random_state(n, 2)
    want, _, _ = oracle.run_program(n, psi, prog.ops)
    got = run_gpu(oracle, prog, state=psi, extra_env={"IQS_B200_MEM": "device"})
    assert np.array_equal(got["state"], want)
