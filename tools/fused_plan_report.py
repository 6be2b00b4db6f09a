"""Host-only report of the schedule iqsb_fused builds for a circuit (no GPU needed).

usage: python tools/fused_plan_report.py [layered|qft|heisenberg] [--n 32] [--layers 1]
Prints, per run: tile size, gates, groups, how many X / CNOT gates were folded into write-back addresses
(absorbed / conditional) and how many run as register moves; plus totals and the sweeps a gate-by-gate
execution would need (one per gate).
"""
import argparse
import collections
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

pkg = g.load_package()
capi, C = pkg.capi, pkg.circuits
CLS = ["general", "real", "diag", "diag1", "anti", "x", "rx", "sqrtx", "sqrty"]


def gates_of(prog):
    import bench

    out = []
    for op in prog.ops:
        k = int(op["kind"])
        if k == C.CGATE1:
            out.append((1, int(op["q0"]), int(op["q1"]), np.array(op["p"][:8])))
        elif k == C.GATE1:
            out.append((0, 0, int(op["q0"]), np.array(op["p"][:8])))
        elif k == C.CX:
            out.append((1, int(op["q0"]), int(op["q1"]), bench.named_matrix(C, k, op["p"])))
        elif k in (C.H, C.RX, C.RY, C.RZ, C.SQRTX, C.SQRTY, C.T):
            out.append((0, 0, int(op["q0"]), bench.named_matrix(C, k, op["p"])))
        # other ops (swaps, reductions) flush the queue in the C++ layer and are not part of a fused batch
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what", nargs="?", default="layered", choices=["layered", "qft", "heisenberg"])
    ap.add_argument("--n", type=int, default=32)
    ap.add_argument("--layers", type=int, default=1)
    a = ap.parse_args()
    prog = {"layered": lambda: C.layered_random(a.n, a.layers), "qft": lambda: C.qft(a.n), "heisenberg": lambda: C.heisenberg_step(a.n, False)}[a.what]()
    gates = gates_of(prog)
    trace, groups = capi.plan_fused_trace(gates, a.n)
    plan = capi.plan_fused_order(gates, a.n)
    inorder = capi.plan_fused_order(gates, a.n, reorder=False)
    print(f"{a.what}, {a.n} qubits: {len(gates)} fusable gates -> {len(plan)} sweeps ({len(inorder)} in program order, {len(gates)} gate by gate), {len(groups)} groups")
    by_run = collections.defaultdict(list)
    for t in trace:
        by_run[t["run"]].append(t)
    nruns = 1 + max(t["run"] for t in trace)
    for r in range(nruns):
        ts = by_run[r]
        tile = sorted({p for t in ts for p in groups[t["group"]]})
        cls = collections.Counter(CLS[t["cls"]] for t in ts)
        x = [t for t in ts if t["cls"] == 5]
        print(f"  run {r}: register positions used {tile}; {len(ts)} gates in {len({t['group'] for t in ts})} groups; classes {dict(cls)}; "
              f"X/CNOT absorbed {sum(1 for t in x if t['trail'] == 2)}, conditional {sum(1 for t in x if t['trail'] == 1)}, as register moves {sum(1 for t in x if t['trail'] == 0)}")


if __name__ == "__main__":
    main()
