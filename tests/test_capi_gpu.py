"""Parity of the CUDA kernels, called through the C ABI (include/iqsb.h), against the CPU oracle.

Bar (BASELINE.json north_star): bit-exact for data movement, <= 1e-12 per amplitude for
arithmetic gates.  The kernels evaluate the reference's exact operation order without FMA
contraction, so most checks below demand bit equality, which is stricter.
"""
import math

import numpy as np
import pytest

from pkg import capi, circuits as C
from progs import random_unitary

pytestmark = pytest.mark.gpu

TOL = 1e-12
X = np.array([0, 0, 1, 0, 1, 0, 0, 0], dtype=np.float64)
HM = np.array([1, 0, 1, 0, 1, 0, -1, 0], dtype=np.float64) / math.sqrt(2.0)


def _rand_m(seed):
    rng = np.random.Generator(np.random.MT19937(seed))
    return np.ascontiguousarray(random_unitary(rng)).ravel().view(np.float64).copy()


@pytest.fixture(scope="module", params=[1, 2, 3, 6, 12])
def small(request, gpu_ctx):
    n = request.param
    st = gpu_ctx.alloc(1 << n)
    yield n, st
    st.free()


def test_gate1_every_position_bit_exact(small, oracle):
    n, st = small
    psi = C.random_state(n, seed=100 + n)
    st.upload(psi)
    ref = psi.copy()
    for pos in range(n):
        m = _rand_m(pos)
        st.gate1(pos, m)
        oracle.gate1(ref, pos, m)
        got = st.download()
        assert np.array_equal(got, ref), f"n={n} pos={pos} maxdiff={np.max(np.abs(got - ref))}"


def test_gate1_subrange(gpu_ctx, oracle):
    n = 10
    st = gpu_ctx.alloc(1 << n)
    psi = C.random_state(n, seed=5)
    st.upload(psi)
    ref = psi.copy()
    m = _rand_m(3)
    for pos, (s, e) in [(0, (64, 192)), (3, (256, 512)), (6, (128, 1024)), (9, (0, 1024))]:
        st.gate1(pos, m, s, e)
        oracle.gate1(ref, pos, m, s, e)
    assert np.array_equal(st.download(), ref)
    with pytest.raises(capi.IqsbError):
        st.gate1(4, m, 8, 1024)  # not aligned to 2^(pos+1)
    with pytest.raises(capi.IqsbError):
        st.gate1(10, m)  # not a local position
    st.free()


def test_cgate1_all_pairs_bit_exact(small, oracle):
    n, st = small
    if n < 2:
        pytest.skip("needs two qubits")
    psi = C.random_state(n, seed=200 + n)
    st.upload(psi)
    ref = psi.copy()
    k = 0
    for c in range(n):
        for t in range(n):
            if c == t:
                continue
            m = _rand_m(1000 + k)
            k += 1
            st.cgate1(c, t, m)
            oracle.cgate1(ref, c, t, m)
    got = st.download()
    assert np.array_equal(got, ref), f"n={n} maxdiff={np.max(np.abs(got - ref))}"


def test_cnot_is_pure_data_movement(gpu_ctx, oracle):
    n = 12
    st = gpu_ctx.alloc(1 << n)
    psi = np.arange(1 << n).astype(np.complex128) + 1j * (np.arange(1 << n) + 0.5)
    st.upload(psi)
    ref = psi.copy()
    for c, t in [(0, 1), (1, 0), (5, 11), (11, 0), (0, 11), (7, 3)]:
        st.cgate1(c, t, X)
        oracle.cgate1(ref, c, t, X)
    assert np.array_equal(st.download(), ref)
    st.free()


def test_swap_family_all_pairs(small, oracle):
    n, st = small
    if n < 2:
        pytest.skip("needs two qubits")
    psi = C.random_state(n, seed=300 + n)
    st.upload(psi)
    ref = psi.copy()
    f = 1 / math.sqrt(2.0)
    mats = [X, np.array([0, 0, 0, 1, 0, 1, 0, 0.0]), np.array([f, 0, 0, f, 0, f, f, 0])]
    k = 0
    for p1 in range(n):
        for p2 in range(p1 + 1, n):
            m = mats[k % 3]
            k += 1
            st.swap2x2(p1, p2, m)
            oracle.swap2x2(ref, p1, p2, m)
    got = st.download()
    assert np.array_equal(got, ref), f"maxdiff={np.max(np.abs(got - ref))}"


def test_swap_golden_vectors(gpu_ctx):
    """unit_test/include/apply_swap_gate_test.hpp:66-106: state j -> j, SWAP(0,1) and SWAP(0,2)."""
    st = gpu_ctx.alloc(8)
    st.upload(np.arange(8).astype(np.complex128))
    st.swap2x2(0, 1, X)
    assert np.array_equal(st.download().real, [0, 2, 1, 3, 4, 6, 5, 7])
    st.upload(np.arange(8).astype(np.complex128))
    st.swap2x2(0, 2, X)
    assert np.array_equal(st.download().real, [0, 4, 2, 6, 1, 5, 3, 7])
    st.free()


def test_swap_equals_three_cnots_bit_exact(gpu_ctx):
    """apply_swap_gate_test.hpp:111-247: SWAP == CX.CX.CX with MaxAbsDiff == 0."""
    n = 10
    a, b = gpu_ctx.alloc(1 << n), gpu_ctx.alloc(1 << n)
    psi = C.random_state(n, seed=9)
    for p1, p2 in [(0, 1), (2, 7), (0, 9), (8, 9)]:
        a.upload(psi)
        b.upload(psi)
        a.swap2x2(p1, p2, X)
        b.cgate1(p1, p2, X)
        b.cgate1(p2, p1, X)
        b.cgate1(p1, p2, X)
        assert a.maxabsdiff(b) == 0.0
        assert a.equal(b)
    a.free()
    b.free()


def test_diag2_scale_phase(small, oracle):
    n, st = small
    if n < 2:
        pytest.skip("needs two qubits")
    rng = np.random.Generator(np.random.MT19937(n))
    psi = C.random_state(n, seed=400 + n)
    st.upload(psi)
    ref = psi.copy()
    for p1 in range(n):
        for p2 in range(n):
            if p1 == p2:
                continue
            d = np.exp(1j * rng.uniform(0, 2 * math.pi, 4))
            st.diag2(p1, p2, d)
            oracle.diag2(ref, p1, p2, d)
    assert np.array_equal(st.download(), ref)
    f = complex(0.3, -0.8)
    st.scale(f)
    oracle.scale(ref, f)
    if n >= 2:
        st.scale(f, 1, 3)  # odd range: width-1 path
        oracle.scale(ref, f, 1, 3)
    assert np.array_equal(st.download(), ref)
    # diag(d0, d1) on one position == generic gate with a diagonal matrix, value for value
    for pos in range(n):
        d0, d1 = complex(math.cos(0.3), -math.sin(0.3)), complex(math.cos(0.3), math.sin(0.3))
        st.phase_by_bit(-1, pos, d0, d1)
        oracle.gate1(ref, pos, np.array([d0.real, d0.imag, 0, 0, 0, 0, d1.real, d1.imag]))
    assert np.array_equal(st.download(), ref)
    for c in range(n):
        for t in range(n):
            if c != t:
                d1 = complex(math.cos(0.7), math.sin(0.7))
                st.phase_by_bit(c, t, 1.0, d1)
                oracle.cgate1(ref, c, t, np.array([1, 0, 0, 0, 0, 0, d1.real, d1.imag]))
    assert np.array_equal(st.download(), ref)


def test_diag2_global_positions(gpu_ctx, oracle):
    """A position >= M takes its bit from glb_start (src/qureg_applydiag.cpp:175-224)."""
    n, M = 8, 6
    full = C.random_state(n, seed=3)
    ref = full.copy()
    d = np.exp(1j * np.array([0.1, 0.2, 0.3, 0.4]))
    oracle.diag2(ref, 7, 2, d)
    oracle.diag2(ref, 1, 6, d)
    oracle.diag2(ref, 6, 7, d)
    st = gpu_ctx.alloc(1 << M)
    out = np.empty_like(full)
    for r in range(1 << (n - M)):
        st.upload(full[r << M : (r + 1) << M])
        st.diag2(7, 2, d, glb_start=r << M)
        st.diag2(1, 6, d, glb_start=r << M)
        st.diag2(6, 7, d, glb_start=r << M)
        out[r << M : (r + 1) << M] = st.download()
    assert np.array_equal(out, ref)
    st.free()


def test_gate2_all_pairs(small, oracle):
    n, st = small
    if n < 2:
        pytest.skip("needs two qubits")
    rng = np.random.Generator(np.random.MT19937(77 + n))
    psi = C.random_state(n, seed=500 + n)
    st.upload(psi)
    ref = psi.copy()
    for ph in range(n):
        for pl in range(n):
            if ph == pl:
                continue
            a = rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4))
            q, _ = np.linalg.qr(a)
            st.gate2(ph, pl, q)
            oracle.gate2(ref, ph, pl, q)
    got = st.download()
    assert np.array_equal(got, ref), f"maxdiff={np.max(np.abs(got - ref))}"


def test_reductions(small, oracle):
    n, st = small
    psi = C.random_state(n, seed=600 + n)
    st.upload(psi)
    assert abs(st.norm2() - oracle.norm2(psi)) <= TOL
    for pos in range(n):
        assert abs(st.prob1(pos) - oracle.prob1(psi, pos)) <= TOL
        for tol in (1e-13, 1e-3, 10.0):
            assert st.any_above(pos, tol) == oracle.any_above(psi, pos, tol)
    if n >= 2:
        for mask in (1, 3, (1 << n) - 1, 1 << (n - 1), 0):
            assert abs(st.parity_expect(mask) - oracle.parity_expect(psi, mask)) <= TOL
        # glb_start contributes the bits of the global qubits
        assert abs(st.parity_expect(1 << n | 1, glb_start=1 << n) - oracle.parity_expect(psi, 1 << n | 1, 1 << n)) <= TOL


def test_two_register_reductions_and_axpy(gpu_ctx, oracle):
    n = 12
    a, b = gpu_ctx.alloc(1 << n), gpu_ctx.alloc(1 << n)
    pa, pb = C.random_state(n, seed=1), C.random_state(n, seed=2)
    a.upload(pa)
    b.upload(pb)
    ov = a.overlap(b)
    assert abs(ov - oracle.overlap(pa, pb)) <= TOL
    assert abs(ov - np.vdot(pb, pa)) <= 1e-12
    f = complex(0.5, 0.25)
    assert abs(a.maxabsdiff(b, f) - oracle.maxabsdiff(pa, pb, f)) <= TOL
    assert abs(a.l2diff(b) - oracle.l2diff(pa, pb)) <= TOL
    assert not a.equal(b)
    a.axpy(b, f)
    ra = oracle.axpy(pa.copy(), pb, f)
    assert np.array_equal(a.download(), ra)
    a.axpy(b, 1.0)
    ra = oracle.axpy(ra, pb, 1.0)
    assert np.array_equal(a.download(), ra)
    b.copy_from(a)
    assert a.equal(b) and a.maxabsdiff(b) == 0.0
    a.free()
    b.free()


def test_collapse_fill_setget(small, oracle):
    n, st = small
    psi = C.random_state(n, seed=700 + n)
    for pos in range(n):
        for val in (0, 1):
            st.upload(psi)
            st.collapse(pos, val)
            assert np.array_equal(st.download(), oracle.collapse(psi.copy(), pos, val))
    st.fill_const(0)
    assert not st.download().any()
    amp = 1.0 / math.sqrt(float(1 << n))
    st.fill_const(amp)
    assert np.array_equal(st.download(), np.full(1 << n, amp, dtype=np.complex128))
    st.set_amp((1 << n) - 1, 1j)
    assert st.get_amp((1 << n) - 1) == 1j
    st.fill_random(12345)
    r1 = st.download()
    st.fill_random(12345)
    assert np.array_equal(r1, st.download())
    assert np.all(np.abs(r1.real) <= 1) and np.all(np.abs(r1.imag) <= 1)
    if n >= 6:
        assert abs(np.mean(r1.real)) < 0.5


def test_permute_local_bit_exact(gpu_ctx):
    n = 12
    st = gpu_ctx.alloc(1 << n)
    psi = C.random_state(n, seed=8)
    rng = np.random.default_rng(0)
    perms = [list(rng.permutation(n)) for _ in range(4)]
    perms.append(list(range(n))[::-1])
    perms.append([0] + list(rng.permutation(n - 1) + 1))  # bit 0 fixed: 32-byte path
    perms.append(list(range(n)))
    idx = np.arange(1 << n, dtype=np.uint64)
    for dst in perms:
        st.upload(psi)
        st.permute_local(dst)
        j = np.zeros_like(idx)
        for b in range(n):
            j |= ((idx >> np.uint64(b)) & np.uint64(1)) << np.uint64(dst[b])
        want = np.empty_like(psi)
        want[j] = psi
        assert np.array_equal(st.download(), want)
    st.free()


def test_permute_local_on_the_bulk_copy_engine(gpu_ctx, monkeypatch):
    """IQS_B200_PERMUTE_BULK=1: tiles travel as 256-byte cp.async.bulk runs with an mbarrier (opt-in,
    measured slower than the default for these short runs); the result is the same, bit for bit."""
    import os

    n = 18
    st = gpu_ctx.alloc(1 << n)
    psi = C.random_state(n, seed=80)
    rng = np.random.default_rng(5)
    idx = np.arange(1 << n, dtype=np.uint64)
    monkeypatch.setenv("IQS_B200_PERMUTE_BULK", "1")
    for dst in [list(rng.permutation(n)) for _ in range(3)] + [list(range(n))[::-1]]:
        st.upload(psi)
        st.permute_local(dst)
        j = np.zeros_like(idx)
        for b in range(n):
            j |= ((idx >> np.uint64(b)) & np.uint64(1)) << np.uint64(dst[b])
        want = np.empty_like(psi)
        want[j] = psi
        assert np.array_equal(st.download(), want)
    st.free()


def test_fused_matches_sequential(gpu_ctx, oracle):
    """Targets and controls anywhere: the batch is cut into tile runs, the result equals the
    gate-by-gate oracle bit for bit."""
    n = 16
    st = gpu_ctx.alloc(1 << n)
    K = st.fused_max_log2tile()
    assert K in (11, 12)
    psi = C.random_state(n, seed=21)
    rng = np.random.Generator(np.random.MT19937(4))
    for lo_only in (True, False):
        st.upload(psi)
        ref = psi.copy()
        gates = []
        for i in range(90):
            m = np.ascontiguousarray(random_unitary(rng)).ravel().view(np.float64).copy()
            t = int(rng.integers(0, K if lo_only else n))
            if i % 3 == 0:
                gates.append((0, 0, t, m))
                oracle.gate1(ref, t, m)
            else:
                c = int(rng.integers(0, n))
                while c == t:
                    c = int(rng.integers(0, n))
                gates.append((1, c, t, m))
                oracle.cgate1(ref, c, t, m)
        runs = capi.plan_fused(gates, n)
        assert runs[0][0] == 0 and runs[-1][1] == len(gates)
        if lo_only:
            assert len(runs) == 1  # every target below the tile exponent: one sweep
        st.fused(gates)
        got = st.download()
        assert np.array_equal(got, ref), f"maxdiff={np.max(np.abs(got - ref))}"
    with pytest.raises(capi.IqsbError):
        st.fused([(0, 0, n, HM)])  # not a local position
    st.free()


def test_fused_contracted_arithmetic_mode(gpu_ctx, oracle):
    """IQSB_ARITH_FMA (opt-in, the reference's IqsNative=ON analogue): same result to rounding, not
    bit for bit; the default mode is restored and stays exact."""
    n = 14
    st = gpu_ctx.alloc(1 << n)
    psi = C.random_state(n, seed=77)
    rng = np.random.Generator(np.random.MT19937(5))
    ref = psi.copy()
    gates = []
    for i in range(60):
        m = np.ascontiguousarray(random_unitary(rng)).ravel().view(np.float64).copy()
        t, c = (int(x) for x in rng.permutation(n)[:2])
        if i % 2:
            gates.append((0, 0, t, m))
            oracle.gate1(ref, t, m)
        else:
            gates.append((1, c, t, m))
            oracle.cgate1(ref, c, t, m)
    assert gpu_ctx.get_arith() == 0
    try:
        gpu_ctx.set_arith(True)
        st.upload(psi)
        st.fused(gates)
        got = st.download()
    finally:
        gpu_ctx.set_arith(False)
    err = np.max(np.abs(got - ref))
    assert err < 1e-14, err  # amplitudes ~ 2^-7, 60 gates: a few ulp
    assert abs(st.norm2() - 1.0) < 1e-13
    st.upload(psi)
    st.fused(gates)
    assert np.array_equal(st.download(), ref)  # exact mode again
    st.free()


def test_small_fused_tiles(gpu_ctx, oracle):
    for n in (2, 3, 5):
        st = gpu_ctx.alloc(1 << n)
        psi = C.random_state(n, seed=n)
        st.upload(psi)
        ref = psi.copy()
        gates = []
        for t in range(n):
            gates.append((0, 0, t, C.G_FIXED))
            oracle.gate1(ref, t, C.G_FIXED)
        gates.append((1, 0, n - 1, X))
        oracle.cgate1(ref, 0, n - 1, X)
        st.fused(gates)
        assert np.array_equal(st.download(), ref)
        st.free()


def test_one_amplitude_registers(gpu_ctx):
    """The reference's default-constructed register holds ONE amplitude (qureg_init.cpp:22-52) and its
    ==, ComputeOverlap, MaxAbsDiff and AmplitudeWiseSum work on it."""
    a, b = gpu_ctx.alloc(1), gpu_ctx.alloc(1)
    a.set_amp(0, 0.6 + 0.8j)
    b.set_amp(0, 0.6 + 0.8j)
    assert a.equal(b)
    assert abs(a.overlap(b) - 1.0) < 1e-15
    b.set_amp(0, 1.0)
    assert not a.equal(b)
    assert abs(a.overlap(b) - np.conj(1.0) * (0.6 + 0.8j)) < 1e-15
    assert abs(a.maxabsdiff(b) - abs(0.6 + 0.8j - 1.0)) < 1e-15
    assert abs(a.l2diff(b) - abs(0.6 + 0.8j - 1.0) ** 2) < 1e-15
    a.axpy(b, 0.5j)
    assert abs(a.get_amp(0) - (0.6 + 0.8j + 0.5j)) < 1e-15
    a.free()
    b.free()


def _pauli_expect_numpy(psi, n, xmask, ymask, zmask):
    """<psi|P|psi> from the definition: P|i> = i^ny (-1)^popcount(i & (y|z)) |i ^ (x|y)>"""
    idx = np.arange(1 << n, dtype=np.uint64)
    f, m = np.uint64(xmask | ymask), np.uint64(ymask | zmask)
    par = np.zeros(1 << n, dtype=np.int64)
    t = idx & m
    for b in range(n):
        par ^= ((t >> np.uint64(b)) & np.uint64(1)).astype(np.int64)
    sign = 1.0 - 2.0 * par
    ny = bin(ymask).count("1")
    val = (1j ** ny) * np.sum(sign * psi * np.conj(psi[idx ^ f]))
    assert abs(val.imag) < 1e-12
    return val.real


@pytest.mark.parametrize("n", [1, 2, 5, 13, 20])
def test_all_marginals_in_one_sweep(gpu_ctx, oracle, n):
    """iqsb_prob_all == [norm, prob1(0), prob1(1), ...] (oracle: the reference's GetProbability sum)"""
    st = gpu_ctx.alloc(1 << n)
    psi = C.random_state(n, seed=50 + n)
    st.upload(psi)
    got = st.prob_all()
    assert got.shape == (n + 1,)
    assert abs(got[0] - 1.0) <= TOL
    for q in range(n):
        assert abs(got[1 + q] - oracle.prob1(psi, q)) <= TOL, q
        if n >= 2 or q > 0:
            assert abs(got[1 + q] - st.prob1(q)) <= TOL
    assert np.array_equal(st.download(), psi)  # nothing written
    st.free()


@pytest.mark.parametrize("n", [1, 2, 6, 14])
def test_pauli_string_expectation_is_read_only_and_matches_definition(gpu_ctx, n):
    st = gpu_ctx.alloc(1 << n)
    psi = C.random_state(n, seed=70 + n)
    st.upload(psi)
    rng = np.random.Generator(np.random.MT19937(n))
    cases = [(1, 0, 0), (0, 1, 0), (0, 0, 1)]
    if n >= 2:
        cases += [(1, 2, 0), (2, 1, 0), (3, 0, 0), (0, 3, 0), (1 << (n - 1), 0, 1), (0, 1 << (n - 1), 1), (1 << (n - 1), 1, 0)]
    for _ in range(12 if n > 2 else 0):
        obs = rng.integers(0, 4, size=n)
        cases.append(tuple(int(sum(1 << q for q in range(n) if obs[q] == k)) for k in (1, 2, 3)))
    for xm, ym, zm in cases:
        if n == 1 and (xm | ym) == 0:
            continue
        got, norm2 = st.pauli_expect(xm, ym, zm, with_norm=True)
        want = _pauli_expect_numpy(psi, n, xm, ym, zm)
        assert abs(got - want) <= TOL, (xm, ym, zm, got, want)
        assert abs(norm2 - 1.0) <= TOL  # sum |a|^2 comes out of the same read
    # a Z factor on a rank bit is taken from glb_start
    if n >= 2:
        hi = 1 << n
        got = st.pauli_expect(1, 0, hi | 2, glb_start=hi)
        assert abs(got + _pauli_expect_numpy(psi, n, 1, 0, 2)) <= TOL
    assert np.array_equal(st.download(), psi)
    with pytest.raises(capi.IqsbError):
        st.pauli_expect(1 << n, 0, 0)  # X on a rank bit is not this kernel's job
    st.free()


def _class_matrices():
    """one matrix per arithmetic class of the fused kernel (kernels_fused.cu: classify)"""
    f = 1.0 / math.sqrt(2.0)
    th = 0.7316
    c, s = math.cos(th / 2), math.sin(th / 2)
    return {
        "general": C.G_FIXED,
        "real": np.array([c, 0, -s, 0, s, 0, c, 0.0]),  # RY
        "hadamard": np.array([f, 0, f, 0, f, 0, -f, 0.0]),
        "rx": np.array([c, 0, 0, -s, 0, -s, c, 0.0]),
        "diag": np.array([c, -s, 0, 0, 0, 0, c, s]),  # RZ
        "diag1": np.array([1, 0, 0, 0, 0, 0, f, f]),  # T
        "z": np.array([1, 0, 0, 0, 0, 0, -1, 0.0]),
        "anti": np.array([0, 0, 0, -1, 0, 1, 0, 0.0]),  # Y
        "x": X,
        "sqrtx": np.array([0.5, 0.5, 0.5, -0.5, 0.5, -0.5, 0.5, 0.5]),
        "sqrty": np.array([0.5, 0.5, -0.5, -0.5, 0.5, 0.5, 0.5, 0.5]),
    }


@pytest.mark.parametrize("n", [4, 5, 7, 11, 12, 15, 18])
def test_fused_gate_classes_and_control_kinds(gpu_ctx, oracle, n):
    """Every matrix class (exact zeros / ones left out of the arithmetic) with no control, a control
    inside the group's register bits, inside the tile and outside the tile; long runs (several
    descriptor batches); tiles smaller than 2^12.  Value-identical to the gate-by-gate oracle."""
    mats = list(_class_matrices().values())
    rng = np.random.Generator(np.random.MT19937(100 + n))
    st = gpu_ctx.alloc(1 << n)
    psi = C.random_state(n, seed=n)
    for trial in range(3):
        st.upload(psi)
        ref = psi.copy()
        gates = []
        ngates = (40, 130, 260)[trial]
        span = (min(n, 4), min(n, 12), n)[trial]  # all targets in a few bits -> deep groups; anywhere -> many runs
        for i in range(ngates):
            m = mats[int(rng.integers(0, len(mats)))]
            t = int(rng.integers(0, span))
            if rng.integers(0, 2):
                gates.append((0, 0, t, m))
                oracle.gate1(ref, t, m)
            else:
                c = int(rng.integers(0, n))
                while c == t:
                    c = int(rng.integers(0, n))
                gates.append((1, c, t, m))
                oracle.cgate1(ref, c, t, m)
        st.fused(gates)
        got = st.download()
        assert np.array_equal(got, ref), f"n={n} trial={trial} maxdiff={np.max(np.abs(got - ref))}"
    st.free()


def test_fused_layered_circuit_reordered_plan_is_bit_exact(gpu_ctx, oracle):
    """The bench circuit (1-qubit gates from the named set + CNOT ladder): CNOTs float into earlier
    runs (iqsb_plan_fused_order), the state equals the program-order oracle bit for bit."""
    n = 20
    prog = C.layered_random(n, 3, seed=5)
    st = gpu_ctx.alloc(1 << n)
    psi = C.random_state(n, seed=9)
    st.upload(psi)
    ref, _, _ = oracle.run_program(n, psi.copy(), prog.ops)
    import bench

    gates = []
    for op in prog.ops:
        m = bench.named_matrix(C, int(op["kind"]), op["p"])
        if op["kind"] == C.CX:
            gates.append((1, int(op["q0"]), int(op["q1"]), m))
        else:
            gates.append((0, 0, int(op["q0"]), m))
    plan = capi.plan_fused_order(gates, n)
    assert [i for run, _ in plan for i in run] != list(range(len(gates)))  # something did move
    assert len(plan) < len(capi.plan_fused(gates, n))
    st.fused(gates)
    assert np.array_equal(st.download(), ref)
    st.free()


def test_float_register_within_float_tolerance(gpu_ctx, oracle):
    """ComplexSP rides the same kernels: checked against the double oracle."""
    n = 10
    st = gpu_ctx.alloc(1 << n, dtype=capi.F32)
    psi = C.random_state(n, seed=31)
    st.upload(psi.astype(np.complex64))
    ref = psi.astype(np.complex64).astype(np.complex128)
    for pos in range(n):
        st.gate1(pos, C.G_FIXED)
        oracle.gate1(ref, pos, C.G_FIXED)
    st.cgate1(0, 5, X)
    oracle.cgate1(ref, 0, 5, X)
    st.swap2x2(1, 8, X)
    oracle.swap2x2(ref, 1, 8, X)
    got = st.download().astype(np.complex128)
    assert np.max(np.abs(got - ref)) < 5e-6
    assert abs(st.norm2() - 1.0) < 1e-5
    st.free()


def test_managed_memory_host_pointer(gpu_ctx, oracle):
    """IQSB_MEM_MANAGED gives the host a real pointer (operator[] / RawState(), SURVEY.md hard part A)."""
    import ctypes

    n = 10
    st = gpu_ctx.alloc(1 << n, mem=capi.MEM_MANAGED)
    psi = C.random_state(n, seed=41)
    st.upload(psi)
    st.gate1(3, C.G_FIXED)
    ref = oracle.gate1(psi.copy(), 3, C.G_FIXED)
    hp = st.L.iqsb_host_ptr(st.h)
    assert hp
    view = np.ctypeslib.as_array(ctypes.cast(hp, ctypes.POINTER(ctypes.c_double)), shape=(2 << n,)).view(np.complex128)
    assert np.array_equal(view, ref)
    view[5] = 2.0 + 1.0j  # host write, then device read
    ref[5] = 2.0 + 1.0j
    st.L.iqsb_prefetch_device(st.h)
    st.gate1(0, C.G_FIXED)
    oracle.gate1(ref, 0, C.G_FIXED)
    assert np.array_equal(st.download(), ref)
    st.free()


def test_large_state_properties(gpu_ctx):
    """Size-independent properties at a size the oracle does not visit (2^26 = 1 GiB):
    unitarity (norm preserved), G then G^dagger restores the state to 1e-12, SWAP round trip is
    bit-exact, and the fused path equals the unfused one bit for bit."""
    n = 26
    a, b = gpu_ctx.alloc(1 << n), gpu_ctx.alloc(1 << n)
    a.fill_random(2024)
    nrm = math.sqrt(a.norm2())
    a.scale(1.0 / nrm)
    b.copy_from(a)
    g = C.G_FIXED.view(np.complex128).reshape(2, 2)
    gd = np.ascontiguousarray(g.conj().T)
    for pos in (0, 1, 5, 13, 25):
        a.gate1(pos, g)
    assert abs(a.norm2() - 1.0) < 1e-12
    for pos in (25, 13, 5, 1, 0):
        a.gate1(pos, gd)
    assert a.maxabsdiff(b) < 1e-12
    a.copy_from(b)
    a.swap2x2(0, 25, X)
    a.swap2x2(3, 17, X)
    a.swap2x2(3, 17, X)
    a.swap2x2(0, 25, X)
    assert a.equal(b)
    gates = [(0, 0, q, C.G_FIXED) for q in range(11)] + [(1, 20, 4, X), (1, 2, 9, HM)] + [(0, 0, q, HM) for q in (25, 13, 19, 22)] + [(1, 3, 24, X)]
    a.fused(gates)
    for kind, c, t, m in gates:
        if kind == 0:
            b.gate1(t, m)
        else:
            b.cgate1(c, t, m)
    assert a.equal(b)
    a.free()
    b.free()


def test_full_size_32_qubits_properties(gpu_ctx):
    """BASELINE configs[1] size (2^32 amplitudes, 64 GiB per register): size-independent properties.
    Skipped when the device cannot hold two registers."""
    free, _ = gpu_ctx.mem_info()
    n = 32
    if free < 2 * 16 * (1 << n) + (4 << 30):
        pytest.skip("needs ~132 GiB of free HBM")
    a, b = gpu_ctx.alloc(1 << n), gpu_ctx.alloc(1 << n)
    a.fill_random(99)
    a.scale(1.0 / math.sqrt(a.norm2()))
    b.copy_from(a)
    g = C.G_FIXED.view(np.complex128).reshape(2, 2)
    gd = np.ascontiguousarray(g.conj().T)
    for pos in (0, 1, 16, 30, 31):
        a.gate1(pos, g)
    a.cgate1(31, 0, X)
    a.cgate1(0, 31, X)
    assert abs(a.norm2() - 1.0) < 1e-12  # unitarity at full size
    a.cgate1(0, 31, X)
    a.cgate1(31, 0, X)
    for pos in (31, 30, 16, 1, 0):
        a.gate1(pos, gd)
    assert a.maxabsdiff(b) < 1e-12  # G then G^dagger restores every amplitude
    a.copy_from(b)
    a.swap2x2(0, 31, X)
    a.swap2x2(7, 30, X)
    a.swap2x2(7, 30, X)
    a.swap2x2(0, 31, X)
    assert a.equal(b)  # data movement round trip is bit exact
    # linearity: P(bit 31 = 1) after H on 31 of the uniform state is 0
    a.fill_const(1.0 / math.sqrt(float(1 << n)))
    a.gate1(31, HM)
    assert abs(a.prob1(31)) < 1e-12 and abs(a.norm2() - 1.0) < 1e-12
    a.free()
    b.free()


def test_float_register_through_the_fused_kernel(gpu_ctx, oracle):
    """ComplexSP through iqsb_fused (same kernel template, 8-byte slots): every matrix class, folded CNOTs,
    controls everywhere, several runs -- against the double oracle within float tolerance."""
    n = 14
    mats = list(_class_matrices().values())
    rng = np.random.Generator(np.random.MT19937(77))
    st = gpu_ctx.alloc(1 << n, dtype=capi.F32)
    psi = C.random_state(n, seed=4)
    st.upload(psi.astype(np.complex64))
    ref = psi.astype(np.complex64).astype(np.complex128)
    gates = []
    for i in range(120):
        m = mats[int(rng.integers(0, len(mats)))]
        t = int(rng.integers(0, n))
        if rng.integers(0, 2):
            gates.append((0, 0, t, m))
            oracle.gate1(ref, t, m)
        else:
            c = int(rng.integers(0, n))
            while c == t:
                c = int(rng.integers(0, n))
            gates.append((1, c, t, m))
            oracle.cgate1(ref, c, t, m)
    st.fused(gates)
    got = st.download().astype(np.complex128)
    assert np.max(np.abs(got - ref)) < 2e-5
    st.free()
