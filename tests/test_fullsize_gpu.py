"""Parity at sizes the oracle restatement does not visit, against INDEPENDENT answers:

 * a closed-form circuit (tests/closed_form.py) at 26 and 32 qubits through the C ABI and at 30
   qubits through the drop-in driver, unfused and fused: >= 10^4 sampled amplitudes, including
   indices with every high bit set, against the analytic product formula at 1e-12;
 * the whole 2^28-amplitude state of a random program against the UNMODIFIED reference compiled
   from its own sources (oracle/_ref/iqs_ref_driver, OpenMP on the host cores);
 * Entropy() / GoogleStats() against that reference build.
"""
import os

import numpy as np
import pytest

from closed_form import ClosedForm
from pkg import circuits as C
from progs import random_program

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "intel-qs_b200", "bin", "iqs_b200_driver")
TOL = 1e-12


@pytest.mark.parametrize("n", [26, 32])
def test_closed_form_amplitudes_capi(gpu_ctx, n):
    """BASELINE configs[1] size on one GPU: every position 0..31 carries its own gate."""
    free, _ = gpu_ctx.mem_info()
    if free < 16 * (1 << n) + (2 << 30):
        pytest.skip(f"needs {16 * (1 << n) >> 30} GiB of free HBM")
    cf = ClosedForm(n)
    samples = cf.samples(10000)
    st = gpu_ctx.alloc(1 << n)
    try:
        for fused in (False, True):
            st.fill_const(0.0)
            st.set_amp(0, 1.0)
            gates = cf.gates()
            if fused:
                st.fused(gates)
            else:
                for kind, c, t, m in gates:
                    if kind == 0:
                        st.gate1(t, m)
                    else:
                        st.cgate1(c, t, m)
            got = [st.get_amp(j) for j in samples]
            err, scale = cf.check(samples, got, TOL)
            print(f"closed form n={n} fused={fused}: {len(samples)} amplitudes, max |d| = {err:.3e} (largest |amp| {scale:.3e})")
            assert abs(st.norm2() - 1.0) < 1e-12
    finally:
        st.free()


@pytest.mark.parametrize("fused", [False, True])
def test_closed_form_amplitudes_dropin_driver_30q(oracle, fused):
    """The same check through iqs::QubitRegister (GetGlobalAmplitude) at 30 qubits."""
    n = 30
    cf = ClosedForm(n, seed=77)
    samples = cf.samples(4000)
    got = oracle.run_driver(DRIVER, cf.program(C, samples, fused=fused), init=1, base_index=0, want_state=False)
    sc = got["scalars"]
    assert sc.size == 2 * len(samples)
    cf.check(samples, sc[0::2] + 1j * sc[1::2], TOL)


def test_full_state_28_qubits_vs_reference_build(oracle, tmp_path):
    """Every one of the 2^28 amplitudes after a random program (all gate kinds, positions up to 27)
    equals what the unmodified reference produces on the host cores."""
    if not oracle.have_ref_driver():
        pytest.skip("oracle/_ref/iqs_ref_driver was not built (no reference tree on the build machine)")
    n = int(os.environ.get("IQS_TEST_FULLSTATE_QUBITS", "28"))
    import shutil

    shm_free = shutil.disk_usage("/dev/shm").free if os.path.isdir("/dev/shm") else 0
    if shm_free < 3 * 16 * (1 << n):
        pytest.skip("not enough /dev/shm for two state files")
    prog = C.Program(n)
    for q in range(n):
        prog.named1(C.RY, q, 0.3 + 0.11 * q)
    prog.extend(random_program(n, 40, seed=31, toffoli=True))
    for q in (0, 1, n - 2, n - 1):
        prog.gate1(q, C.G_FIXED)
    prog.named2(C.CX, n - 1, 0).named2(C.CX, 0, n - 1).named2(C.SWAP, 1, n - 1)
    ref = oracle.run_reference(prog, init=1, base_index=5)
    got = oracle.run_driver(DRIVER, prog, init=1, base_index=5)
    a, b = got["state"], ref["state"]
    assert a.size == b.size == 1 << n
    worst = 0.0
    step = 1 << 24
    for i in range(0, a.size, step):
        worst = max(worst, float(np.max(np.abs(a[i : i + step] - b[i : i + step]))))
    print(f"2^{n} amplitudes vs the reference build: max |d| = {worst:.3e}, GPU {got['seconds']:.3f} s, reference {ref['seconds']:.3f} s")
    assert worst <= TOL
    assert np.array_equal(got["map"], ref["map"])


def test_entropy_and_google_stats_vs_reference_build(oracle):
    """Entropy() / GoogleStats() (reference src/qureg_utils.cpp:305-450): 12 qubits, against the
    reference build when present and against the oracle restatement always."""
    n = 12
    prog = random_program(n, 150, seed=9)
    prog.entropy().google_stats()
    prog.named1(C.H, 3).collapse(3, 1).normalize().entropy().google_stats()  # zero amplitudes: the p != 0 branch
    psi = C.random_state(n, 9)
    got = oracle.run_driver(DRIVER, prog, state=psi, want_state=False)["scalars"]
    _, want, _ = oracle.run_program(n, psi, prog.ops)
    assert got.size == want.size == 24

    def close(x, y):
        return np.all(np.abs(x - y) <= 1e-12 * np.maximum(1.0, np.abs(y)))

    assert close(got, want), (got, want)
    if oracle.have_ref_driver():
        ref = oracle.run_reference(prog, state=psi, want_state=False, threads=1)["scalars"]
        assert close(got, ref), (got, ref)
