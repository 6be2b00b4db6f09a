"""Per-kernel device timing through the C ABI (CUDA events on the engine's stream).

usage: python tools/kbench.py [--n 32] [--reps 5] [--mem device|managed] [--out gpurun_out/kbench.json]
Prints one line per (op, position): ms per launch, algorithmic GB/s and the
fraction of the measured HBM peak (MEASURED_PEAKS.json).
"""
import argparse
import json
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

pkg = g.load_package()
capi, C = pkg.capi, pkg.circuits


def peak_gbs():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "measured"
    except Exception:
        return 6650.0, "fallback"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=32)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--mem", default="device")
    ap.add_argument("--out", default=None)
    ap.add_argument("--arith", default="exact", choices=["exact", "fma"], help="arithmetic of the fused kernel (iqsb_set_arith)")
    ap.add_argument("--ops", default="gate1,cgate1,swap,diag,phase,prob,norm,parity,fused,gate2,collapse,permute")
    a = ap.parse_args()
    n = a.n
    L = 1 << n
    peak, how = peak_gbs()
    ctx = capi.Context()
    ctx.set_arith(a.arith == "fma")
    st = ctx.alloc(L, mem=capi.MEM_MANAGED if a.mem == "managed" else capi.MEM_DEVICE)
    st.fill_random(1)
    st.scale(1.0 / math.sqrt(st.norm2()))
    X = np.array([0, 0, 1, 0, 1, 0, 0, 0.0])
    rows = []

    def timeit(name, pos, fn, bytes_algo):
        fn()  # warm
        ctx.sync()
        best = 1e30
        tot = 0.0
        for _ in range(a.reps):
            ctx.timer_start()
            fn()
            ms = ctx.timer_stop()
            best = min(best, ms)
            tot += ms
        gbs = bytes_algo / (best * 1e-3) / 1e9
        rows.append({"op": name, "pos": pos, "ms_best": best, "ms_mean": tot / a.reps, "algo_gbs": gbs, "frac": gbs / peak})
        print(f"{name:10s} pos={str(pos):8s} {best:9.3f} ms  {gbs:8.1f} GB/s  {100 * gbs / peak:5.1f}% of {how} {peak:.0f}", flush=True)

    ops = a.ops.split(",")
    if "gate1" in ops:
        for pos in range(n):
            timeit("gate1", pos, lambda pos=pos: st.gate1(pos, C.G_FIXED), 32.0 * L)
    if "cgate1" in ops:
        for c, t in [(0, 1), (1, 0), (1, 2), (2, 1), (5, 6), (6, 5), (0, n - 1), (n - 1, 0), (n - 2, n - 1), (n - 1, n - 2), (10, 20), (20, 10)]:
            timeit("cgate1", (c, t), lambda c=c, t=t: st.cgate1(c, t, X), 16.0 * L)
    if "swap" in ops:
        for p1, p2 in [(0, 1), (0, n - 1), (1, 2), (3, 17), (n - 2, n - 1)]:
            timeit("swap", (p1, p2), lambda p1=p1, p2=p2: st.swap2x2(p1, p2, X), 16.0 * L)
    if "diag" in ops:
        d = np.exp(1j * np.array([0.1, 0.2, 0.3, 0.4]))
        for p1, p2 in [(0, 1), (3, 17), (n - 1, 0)]:
            timeit("diag2", (p1, p2), lambda p1=p1, p2=p2: st.diag2(p1, p2, d), 32.0 * L)
    if "phase" in ops:
        for pos in (0, 1, 10, n - 1):
            timeit("phaseT", pos, lambda pos=pos: st.phase_by_bit(-1, pos, 1.0, complex(math.cos(0.7), math.sin(0.7))), 16.0 * L)
        for c, t in [(0, 1), (5, 20), (n - 1, 3)]:
            timeit("cphase", (c, t), lambda c=c, t=t: st.phase_by_bit(c, t, 1.0, complex(math.cos(0.7), math.sin(0.7))), 8.0 * L)
    if "prob" in ops:
        for pos in (0, 1, 10, n - 1):
            timeit("prob1", pos, lambda pos=pos: st.prob1(pos), 8.0 * L)
    if "prob" in ops:
        timeit("prob_all", "-", lambda: st.prob_all(), 16.0 * L)
    if "pauli" in ops or "parity" in ops:
        for name, (xm, ym, zm) in (("X0", (1, 0, 0)), ("X5Y9Z1", (1 << 5, 1 << 9, 2)), ("XtopY0", (1 << (n - 1), 1, 0)), ("X0..3", (15, 0, 0))):
            timeit("pauli", name, lambda xm=xm, ym=ym, zm=zm: st.pauli_expect(xm, ym, zm), 16.0 * L)
    if "norm" in ops:
        timeit("norm2", "-", lambda: st.norm2(), 16.0 * L)
    if "parity" in ops:
        timeit("parity", "-", lambda: st.parity_expect((1 << n) - 1), 16.0 * L)
    if "fused" in ops:
        H = np.array([1, 0, 1, 0, 1, 0, -1, 0.0]) / math.sqrt(2)
        for ng in (1, 4, 8, 16, 32):
            gates = [(0, 0, i % 11, C.G_FIXED) for i in range(ng)]
            timeit(f"fused{ng}", "-", lambda gates=gates: st.fused(gates), 32.0 * L)
        # one run whose tile is built from positions scattered over the whole index (4 low + 7 high)
        hi = [n - 1 - 3 * i for i in range(7)]
        for ng in (1, 7, 28):
            gates = [(0, 0, hi[i % 7], C.G_FIXED) for i in range(ng)]
            timeit(f"fused_hi{ng}", "-", lambda gates=gates: st.fused(gates), 32.0 * L)
        # a layer of one-qubit gates on every qubit followed by a CX chain (several runs)
        gates = [(0, 0, q, C.G_FIXED) for q in range(n)] + [(1, q, q + 1, X) for q in range(0, n - 1, 2)]
        runs = len(capi.plan_fused_order(gates, n))
        timeit(f"fused_layer{len(gates)}g{runs}r", "-", lambda gates=gates: st.fused(gates), 32.0 * L * runs)
        # in-tile cost per matrix class: 12 / 48 gates on the 12 lowest positions (one run)
        f, c_, s_ = 1 / math.sqrt(2), math.cos(0.4), math.sin(0.4)
        classes = {"general": C.G_FIXED, "real": H, "rx": np.array([c_, 0, 0, -s_, 0, -s_, c_, 0.0]), "diag": np.array([c_, -s_, 0, 0, 0, 0, c_, s_]),
                   "diag1": np.array([1, 0, 0, 0, 0, 0, f, f]), "x": X}
        for cname, m in classes.items():
            for ng in (12, 48):
                gates = [(0, 0, i % 12, m) for i in range(ng)]
                timeit(f"fused{ng}_{cname}", "-", lambda gates=gates: st.fused(gates), 32.0 * L)
        # the layers bench.py times (named 1-qubit gates + CNOT ladder), one call per layer
        import bench

        for li, layer in enumerate(bench.build_layers(C, n, 2)):
            gates = [((1, int(op["q0"]), int(op["q1"])) if op["kind"] == C.CX else (0, 0, int(op["q0"]))) + (bench.named_matrix(C, int(op["kind"]), op["p"]),) for op in layer]
            runs = len(capi.plan_fused_order(gates, n))
            timeit(f"bench_layer{li}_{len(gates)}g{runs}r", "-", lambda gates=gates: st.fused(gates), 32.0 * L * runs)
    if "fusedprof" in ops:  # the profiling target: one bench layer (named 1-qubit gates + CNOT ladder) as one fused call
        import bench

        layer = bench.build_layers(C, n, 1)[0]
        gates = [((1, int(op["q0"]), int(op["q1"])) if op["kind"] == C.CX else (0, 0, int(op["q0"]))) + (bench.named_matrix(C, int(op["kind"]), op["p"]),) for op in layer]
        runs = len(capi.plan_fused_order(gates, n))
        timeit(f"bench_layer0_{len(gates)}g{runs}r", "-", lambda gates=gates: st.fused(gates), 32.0 * L * runs)
    if "fusedx" in ops:  # profiling target: 12 X gates on the 12 lowest positions = 4 groups, no arithmetic
        gates = [(0, 0, i % 12, X) for i in range(12)]
        timeit("fused12_x", "-", lambda gates=gates: st.fused(gates), 32.0 * L)
    if "gate2" in ops:
        rng = np.random.default_rng(0)
        q, _ = np.linalg.qr(rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4)))
        for ph, pl in [(1, 0), (0, 1), (5, 3), (n - 1, 2)]:
            timeit("gate2", (ph, pl), lambda ph=ph, pl=pl: st.gate2(ph, pl, q), 32.0 * L)
    if "permute" in ops:
        rev = list(range(n))[::-1]
        rot = [(b + 1) % n for b in range(n)]
        low = list(range(n)); low[1], low[n - 2] = low[n - 2], low[1]
        for name, perm in (("reverse", rev), ("rotate1", rot), ("swap(1,n-2)", low)):
            timeit("permute", name, lambda perm=perm: st.permute_local(perm), 32.0 * L)
    if "collapse" in ops:
        for pos in (0, 5, n - 1):
            timeit("collapse", pos, lambda pos=pos: st.collapse(pos, 1), 8.0 * L)
    if a.out:
        os.makedirs(os.path.dirname(a.out), exist_ok=True)
        json.dump({"n": n, "mem": a.mem, "peak_gbs": peak, "peak_kind": how, "rows": rows}, open(a.out, "w"), indent=1)
    st.free()
    ctx.close()


if __name__ == "__main__":
    main()
