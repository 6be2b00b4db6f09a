#!/usr/bin/env python
"""fuzz_fused_model.py -- long-running companion of tests/test_fused_plan.py (no GPU needed).

Random circuits (4..15 qubits, 1..260 gates over every matrix class, random share of controls and of
X / CNOT gates, random target span) are planned by iqsb_plan_fused_dump, the raw descriptors are executed by
the numpy model of k_fused (tests/fused_model.py) and the result must equal the gate-by-gate oracle bit for
bit, with and without reordering.

usage: python tools/fuzz_fused_model.py FIRST_SEED COUNT
Round 2: seeds 200000..203199 (3200 circuits x 2 modes): 0 mismatches.
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import numpy as np
import conftest  # noqa
from pkg import capi, circuits as C
import fused_model
import importlib.util
spec = importlib.util.spec_from_file_location("iqs_oracle", os.path.join(ROOT, "oracle", "oracle.py"))
oracle = importlib.util.module_from_spec(spec); spec.loader.exec_module(oracle)
import test_fused_plan as T
mats = T._model_matrices()
seed0 = int(sys.argv[1]); count = int(sys.argv[2])
t0 = time.time(); bad = 0
for s in range(seed0, seed0 + count):
    rng = np.random.Generator(np.random.MT19937(s))
    n = int(rng.integers(4, 16))
    ngates = int(rng.integers(1, 260))
    span = int(rng.integers(1, n + 1))
    pctrl = rng.random()
    px = rng.random() * 0.7
    psi = C.random_state(n, seed=s)
    gates = []
    for i in range(ngates):
        m = mats[4] if rng.random() < px else mats[int(rng.integers(0, len(mats)))]
        t = int(rng.integers(0, span))
        if rng.random() > pctrl or n == 1:
            gates.append((0, 0, t, m))
        else:
            c = int(rng.integers(0, n))
            while c == t: c = int(rng.integers(0, n))
            gates.append((1, c, t, m))
    want = T._oracle_apply(oracle, psi, gates)
    for reorder in (True, False):
        try:
            got = fused_model.run(psi, capi.plan_fused_dump(gates, n, reorder), n)
            ok = np.array_equal(got, want)
        except Exception as e:
            ok = False; print("EXC", s, n, ngates, reorder, repr(e)[:200], flush=True)
        if not ok:
            bad += 1; print("MISMATCH seed", s, "n", n, "gates", ngates, "span", span, "reorder", reorder, flush=True)
print("done", seed0, count, "bad", bad, "sec", round(time.time() - t0), flush=True)
