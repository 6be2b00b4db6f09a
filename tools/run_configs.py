"""Time the BASELINE.json configurations that are not the bench line, through the drop-in driver
(iqs::QubitRegister API on the GPU), and print one JSON line per run.

  python tools/run_configs.py qft --n 20                       # config 1 (plumbing)
  python tools/run_configs.py heisenberg --n 32 [--fusion 11]  # config 5
  python tools/run_configs.py qft --n 34 --ranks 2             # config 3 (uses tools/iqsrun)
  python tools/run_configs.py reorder --n 33 --ranks 2         # config 4 circuit (gates on the top qubits, trivial vs reversed order)
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

pkg = g.load_package()
C = pkg.circuits
orc = g.load_oracle()  # only its run_driver helper (subprocess plumbing) is used here
DRIVER = os.path.join(ROOT, "intel-qs_b200", "bin", "iqs_b200_driver")
IQSRUN = os.path.join(ROOT, "tools", "iqsrun")


def classify(prog, n, M):
    """count gates by where their qubits live (identity map)"""
    k = {"local": 0, "global_diag": 0, "global_exchange": 0}
    for op in prog.ops:
        kind = int(op["kind"])
        if kind >= 50:
            continue
        qs = [int(op["q0"])] + ([int(op["q1"])] if kind in (C.CGATE1, C.SWAP, C.CX, C.CPHASE, C.CZ) else [])
        if all(q < M for q in qs):
            k["local"] += 1
        elif kind == C.CGATE1 and op["p"][2] == 0 and op["p"][3] == 0 and op["p"][4] == 0 and op["p"][5] == 0:
            k["global_diag"] += 1
        else:
            k["global_exchange"] += 1
    return k


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["qft", "heisenberg", "reorder", "sweep", "layered"])
    ap.add_argument("--n", type=int, default=20)
    ap.add_argument("--ranks", type=int, default=1)
    ap.add_argument("--fusion", type=int, default=0)
    ap.add_argument("--repeat", type=int, default=1)
    a = ap.parse_args()
    n = a.n
    M = n - int(np.log2(a.ranks))
    launcher = [sys.executable, IQSRUN, "-n", str(a.ranks), "--timeout", "1200"] if a.ranks > 1 else None
    runs = []
    if a.what == "qft":
        p = C.qft(n)
        if a.fusion:
            p = C.Program(n).mode(C.FUSION_ON, a.fusion).extend(p).mode(C.FUSION_OFF)
        runs.append((f"qft{' fused' if a.fusion else ''}", p, 2, 0))
    elif a.what == "sweep":
        runs.append(("basic_code_for_scaling sweep", C.scaling_sweep(n), 1, 0))
    elif a.what == "layered":
        p = C.layered_random(n, 5)
        if a.fusion:
            p = C.Program(n).mode(C.FUSION_ON, a.fusion).extend(p).mode(C.FUSION_OFF)
        runs.append((f"layered random circuit, 5 layers{' fused' if a.fusion else ''}", p, 2, 0))
    elif a.what == "heisenberg":
        p = C.heisenberg_step(n)
        if a.fusion:
            p = C.Program(n).mode(C.FUSION_ON, a.fusion).extend(p).mode(C.FUSION_OFF)
        runs.append((f"heisenberg step{' fused' if a.fusion else ''}", p, 1, 1))
    else:
        # examples/communication_reduction_via_qubit_reordering.cpp:99-132: H, X, Y, Z on the top 10 qubits,
        # once in the trivial order and once with the qubit order reversed (PermuteQubits included in the time)
        top = list(range(n - 10, n))
        body = C.Program(n)
        for _ in range(2):
            for q in top:
                body.named1(C.H, q).named1(C.X, q).named1(C.Y, q).named1(C.Z, q)
        def wrap(p):  # --fusion applies to this config too
            return C.Program(n).mode(C.FUSION_ON, a.fusion).extend(p).mode(C.FUSION_OFF) if a.fusion else p

        tag = " fused" if a.fusion else ""
        runs.append(("trivial order" + tag, wrap(body), 1, 0))
        rev = C.Program(n).permute([n - 1 - q for q in range(n)]).extend(wrap(body)).permute(list(range(n)))
        runs.append(("reversed order (2 PermuteQubits included)" + tag, rev, 1, 0))
    for name, prog, init, base in runs:
        r = orc.run_driver(DRIVER, prog, init=init, base_index=base, want_state=False, launcher=launcher, repeat=a.repeat)
        gates = prog.count_gates() * a.repeat
        out = {"config": a.what, "run": name, "qubits": n, "ranks": a.ranks, "local_qubits": M, "gates": gates, "seconds": r["seconds"],
               "gates_per_s": gates / r["seconds"], "where": classify(prog, n, M), "scalars_head": [float(x) for x in r["scalars"][:4]]}
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
