"""CPU: how iqsb_fused cuts a batch of gates into shared-memory tile runs (iqsb_plan_fused, pure host code)."""
import numpy as np
import pytest

from pkg import capi

H = np.array([1, 0, 1, 0, 1, 0, -1, 0.0]) / np.sqrt(2)
X = np.array([0, 0, 1, 0, 1, 0, 0, 0.0])
K = len(capi.plan_fused([(0, 0, 0, H)], 40)[0][2])  # tile exponent of the faster CTA shape (2^11 amplitudes, 128 threads)
KBIG = K + 1  # tiles of 2^12 amplitudes (256 threads): taken when they save a run


def tile_sizes_ok(tiles, M):
    sizes = {len(t) for t in tiles}
    return len(sizes) == 1 and sizes <= {min(K, M), min(KBIG, M)}


def check_cover(gates, runs, M):
    assert runs[0][0] == 0 and runs[-1][1] == len(gates)
    assert tile_sizes_ok([tile for _, _, tile in runs], M)
    for (first, last, tile), nxt in zip(runs, runs[1:] + [None]):
        assert first < last and tile == sorted(set(tile))
        assert tile[: min(4, M)] == list(range(min(4, M)))  # low positions always inside: 256-byte runs
        for k in range(first, last):
            assert gates[k][2] in tile  # every target of the run is in its tile
        if nxt is not None:
            assert nxt[0] == last


def test_layer_of_32_one_qubit_gates_needs_four_sweeps():
    gates = [(0, 0, q, H) for q in range(32)]
    runs = capi.plan_fused(gates, 32)
    check_cover(gates, runs, 32)
    assert len(runs) == 1 + -(-(32 - K) // (K - 4))  # K = 11: qubits 0-10, 11-17, 18-24, 25-31
    assert runs[0][2] == list(range(K))


def test_layered_circuit_runs():
    rng = np.random.default_rng(1)
    gates = []
    for layer in range(3):
        gates += [(0, 0, q, H) for q in range(32)]
        gates += [(1, q, q + 1, X) for q in range(layer % 2, 31, 2)]
    runs = capi.plan_fused(gates, 32)
    check_cover(gates, runs, 32)
    assert len(runs) <= 24  # 144 gates in at most 24 sweeps


def test_random_batches_are_covered_in_order():
    rng = np.random.default_rng(7)
    for M in (1, 2, 3, 7, 12, 20, 33):
        gates = []
        for i in range(200):
            t = int(rng.integers(0, M))
            if M > 1 and i % 2:
                c = int(rng.integers(0, M))
                while c == t:
                    c = int(rng.integers(0, M))
                gates.append((1, c, t, X))
            else:
                gates.append((0, 0, t, H))
        check_cover(gates, capi.plan_fused(gates, M), M)


def test_controls_prefer_tile_slots():
    gates = [(1, 30, 5, X), (1, 29, 6, X)]
    (first, last, tile), = capi.plan_fused(gates, 32)
    assert 30 in tile and 29 in tile  # spare slots are given to the controls of the run


# ---- the plan iqsb_fused executes: exact X / CNOT gates may move ahead of gates on other qubits ----
def _qubits(g):
    return {g[2]} | ({g[1]} if g[0] == 1 else set())


def _is_perm(g):
    return np.array_equal(np.asarray(g[3], dtype=float), X)


def check_order(gates, plan, M):
    order = [i for run, _ in plan for i in run]
    assert sorted(order) == list(range(len(gates)))  # every gate exactly once
    when = {i: k for k, i in enumerate(order)}
    assert tile_sizes_ok([tile for _, tile in plan], M)
    for run, tile in plan:
        assert tile == sorted(set(tile))
        for i in run:
            assert gates[i][2] in tile
    for i in range(len(gates)):
        for j in range(i + 1, len(gates)):
            if when[j] < when[i]:  # j overtook i: allowed only if that is exact
                assert not (_qubits(gates[i]) & _qubits(gates[j])), (i, j)
                assert _is_perm(gates[i]) or _is_perm(gates[j]), (i, j)


def test_cnots_float_into_the_run_of_their_qubits():
    """a layer of 32 one-qubit gates + 16 CNOTs: 4 sweeps (6 in program order), for both CNOT parities"""
    for parity in (0, 1):
        gates = [(0, 0, q, H) for q in range(32)] + [(1, q, q + 1, X) for q in range(parity, 31, 2)]
        plan = capi.plan_fused_order(gates, 32)
        check_order(gates, plan, 32)
        assert len(plan) == 4
        assert len(capi.plan_fused_order(gates, 32, reorder=False)) == len(capi.plan_fused(gates, 32)) == 6


def test_without_reorder_the_plan_is_the_program_order():
    rng = np.random.default_rng(3)
    gates = [(int(rng.integers(0, 2)), int(c), int(t), X if rng.integers(0, 2) else H) for c, t in (rng.permutation(20)[:2] for _ in range(150))]
    plan = capi.plan_fused_order(gates, 20, reorder=False)
    assert [i for run, _ in plan for i in run] == list(range(len(gates)))
    assert [(len(run), tile) for run, tile in plan] == [(last - first, tile) for first, last, tile in capi.plan_fused(gates, 20)]


def test_reordered_plans_respect_exact_commutation():
    rng = np.random.default_rng(11)
    for M in (5, 13, 20, 33):
        gates = []
        for i in range(160):
            c, t = (int(x) for x in rng.permutation(M)[:2])
            kind = int(rng.integers(0, 2))
            gates.append((kind, c, t, X if rng.integers(0, 3) else H))
        plan = capi.plan_fused_order(gates, M)
        check_order(gates, plan, M)
        assert len(plan) <= len(capi.plan_fused(gates, M))


# ---- the full schedule: runs -> groups (three register bits) -> gates with class and control kind ----
G = np.array([0.5922, 0.4596, -0.0387, -0.6608, -0.1305, 0.6490, 0.4966, 0.5621])  # a general matrix
T = np.array([1, 0, 0, 0, 0, 0, np.cos(np.pi / 4), np.sin(np.pi / 4)])
SX = np.array([0.5, 0.5, 0.5, -0.5, 0.5, -0.5, 0.5, 0.5])
CLS = {"general": 0, "real": 1, "diag": 2, "diag1": 3, "anti": 4, "x": 5, "rx": 6, "sqrtx": 7, "sqrty": 8}


def check_trace(gates, M, reorder=True):
    trace, groups = capi.plan_fused_trace(gates, M, reorder)
    assert sorted(t["gate"] for t in trace) == list(range(len(gates)))
    when = {t["gate"]: k for k, t in enumerate(trace)}
    for a, b in zip(trace, trace[1:]):
        assert (b["run"], b["group"]) >= (a["run"], a["group"])  # runs and groups are executed in order
    for t in trace:
        g = gates[t["gate"]]
        regs = groups[t["group"]]
        assert len(regs) == 3 and len(set(regs)) == 3
        assert regs[t["tbit"]] == g[2]  # the target is the register bit the kernel is told
        if g[0] == 1:
            assert t["ckind"] in (1, 2, 3)
            # only an exact X takes its control among the register bits (a move between registers);
            # the arithmetic classes are straight-line code over all register pairs
            if t["ckind"] == 1:
                assert t["cls"] == CLS["x"] and regs[t["c"]] == g[1]
            else:
                assert g[1] not in regs
        else:
            assert t["ckind"] == 0
    info = {t["gate"]: t for t in trace}
    for t in trace:
        if t["trail"]:
            assert t["cls"] == CLS["x"] and reorder  # only exact X / CNOT are folded into the write-back addresses
            assert (t["trail"] == 1) == (t["ckind"] in (2, 3))
    for i in range(len(gates)):
        for j in range(i + 1, len(gates)):
            if when[j] < when[i]:
                a, b = info[i], info[j]
                if a["trail"] and b["trail"] and a["group"] == b["group"]:
                    continue  # both folded into the same write-back: composed in program order by construction
                assert not (_qubits(gates[i]) & _qubits(gates[j])), (i, j)
                assert _is_perm(gates[i]) or _is_perm(gates[j]), (i, j)
    return trace, groups


def test_schedule_classes_and_controls():
    gates = [(0, 0, 0, G), (0, 0, 1, H), (0, 0, 2, T), (0, 0, 3, SX), (1, 0, 1, X), (1, 2, 3, X), (1, 5, 4, G), (1, 4, 5, H), (0, 0, 6, X)]
    trace, groups = check_trace(gates, 20)
    cls = {t["gate"]: t["cls"] for t in trace}
    assert [cls[i] for i in range(4)] == [CLS["general"], CLS["real"], CLS["diag1"], CLS["sqrtx"]]
    assert cls[4] == CLS["x"] and cls[8] == CLS["x"]
    kinds = {t["gate"]: t["ckind"] for t in trace}
    assert kinds[4] == 1  # CNOT(0,1): both in the first group's registers
    assert kinds[6] == 2 and kinds[7] == 2  # controlled arithmetic gates: control on a thread bit
    # controlled-G(5 -> 4) and controlled-H(4 -> 5) cannot share a group: each one's control is the other's target
    g_of = {t["gate"]: t["group"] for t in trace}
    assert g_of[6] != g_of[7]


def test_schedule_of_a_bench_layer_absorbs_the_cnots():
    """32 one-qubit gates + 16 CNOTs: 4 runs; in the first run (qubits 0..11) the 6 CNOTs join the 4
    groups of the one-qubit gates instead of opening their own."""
    for parity in (0, 1):
        gates = [(0, 0, q, (G, H, T, SX)[q % 4]) for q in range(32)] + [(1, q, q + 1, X) for q in range(parity, 31, 2)]
        trace, groups = check_trace(gates, 32)
        assert max(t["run"] for t in trace) == 3
        first_run = [t for t in trace if t["run"] == 0]
        assert len({t["group"] for t in first_run}) == 4
        assert sum(1 for t in first_run if t["cls"] == CLS["x"]) >= 5
        assert all(t["trail"] for t in first_run if t["cls"] == CLS["x"])  # the CNOTs cost no instruction on the data
        assert len(groups) <= 14


def test_random_schedules():
    rng = np.random.default_rng(23)
    mats = [G, H, T, SX, X, X]
    for M in (4, 6, 12, 13, 33):
        gates = []
        for i in range(180):
            c, t = (int(x) for x in rng.permutation(M)[:2])
            gates.append((int(rng.integers(0, 2)), c, t, mats[int(rng.integers(0, len(mats)))]))
        check_trace(gates, M)
        check_trace(gates, M, reorder=False)


def test_tile_size_is_the_small_one_unless_it_costs_a_run():
    # 12 distinct targets: one run of 2^12-amplitude tiles, two of 2^11 -> the big tile
    gates = [(0, 0, q, H) for q in range(12)]
    (first, last, tile), = capi.plan_fused(gates, 32)
    assert len(tile) == KBIG
    # 32 targets: four runs either way -> the small tile (faster CTA shape)
    gates = [(0, 0, q, H) for q in range(32)]
    runs = capi.plan_fused(gates, 32)
    assert len(runs) == 4 and all(len(t) == K for _, _, t in runs)
    # a register smaller than the small tile is one tile
    (first, last, tile), = capi.plan_fused([(0, 0, 3, H)], 9)
    assert len(tile) == 9


# ---- the schedule executed by a CPU model of the kernel (tests/fused_model.py) == the gate-by-gate oracle ----
def _oracle_apply(oracle, psi, gates):
    ref = psi.copy()
    for kind, c, t, m in gates:
        if kind == 0:
            oracle.gate1(ref, t, np.ascontiguousarray(m, dtype=np.float64))
        else:
            oracle.cgate1(ref, c, t, np.ascontiguousarray(m, dtype=np.float64))
    return ref


def _model_matrices():
    f, th = 1 / np.sqrt(2), 0.7316
    c, s_ = np.cos(th / 2), np.sin(th / 2)
    return [G, H, T, SX, X, X, np.array([0.5, 0.5, -0.5, -0.5, 0.5, 0.5, 0.5, 0.5]),  # general, real, diag(1,d), sqrt X, X, X, sqrt Y
            np.array([c, 0, 0, -s_, 0, -s_, c, 0.0]), np.array([c, -s_, 0, 0, 0, 0, c, s_]),  # RX, RZ
            np.array([0, 0, 0, -1, 0, 1, 0, 0.0]), np.array([c, 0, -s_, 0, s_, 0, c, 0.0])]  # Y, RY


@pytest.mark.parametrize("n", [4, 5, 8, 11, 12, 14])
def test_cpu_model_of_the_kernel_reproduces_the_oracle(oracle, n):
    """Random circuits over every matrix class, with controls in registers / on thread bits / outside the
    tile and many X / CNOT gates (floated, folded into write-back addresses, conditional offsets): the
    descriptors iqsb_fused builds, executed by the CPU model of the kernel, give the oracle's state bit
    for bit -- with and without reordering."""
    import fused_model
    from pkg import circuits as C

    rng = np.random.Generator(np.random.MT19937(900 + n))
    mats = _model_matrices()
    psi = C.random_state(n, seed=n)
    for trial, (ngates, span) in enumerate(((60, min(n, 4)), (150, n), (150, min(n, 7)))):
        gates = []
        for i in range(ngates):
            m = mats[int(rng.integers(0, len(mats)))]
            t = int(rng.integers(0, span))
            if rng.integers(0, 2):
                gates.append((0, 0, t, m))
            else:
                c = int(rng.integers(0, n))
                while c == t:
                    c = int(rng.integers(0, n))
                gates.append((1, c, t, m))
        want = _oracle_apply(oracle, psi, gates)
        for reorder in (True, False):
            got = fused_model.run(psi, capi.plan_fused_dump(gates, n, reorder), n)
            assert np.array_equal(got, want), (n, trial, reorder, np.max(np.abs(got - want)))


def test_cpu_model_on_bench_layers(oracle):
    """the layered circuit of bench.py at 16 qubits, three layers in one call and layer by layer"""
    import bench
    import fused_model
    from pkg import circuits as C

    n = 16
    layers = bench.build_layers(C, n, 3)
    psi = C.random_state(n, seed=3)
    allg = []
    for layer in layers:
        allg += [((1, int(op["q0"]), int(op["q1"])) if op["kind"] == C.CX else (0, 0, int(op["q0"]))) + (bench.named_matrix(C, int(op["kind"]), op["p"]),) for op in layer]
    want = _oracle_apply(oracle, psi, allg)
    got = fused_model.run(psi, capi.plan_fused_dump(allg, n), n)
    assert np.array_equal(got, want)
    trace, _ = capi.plan_fused_trace(allg, n)
    assert any(t["trail"] == 2 for t in trace) and any(t["trail"] == 1 for t in trace)  # both kinds of folded CNOTs occur
