"""bench.py --impl reference on the CPU (the arm never touches a GPU): the JSON contract of the line, and under a
2-process torchrun launch only rank 0 works and prints.  The register is shrunk with IQS_BENCH_CPU_QUBITS so
the test takes seconds; the arm itself sizes the sample from the host's RAM."""
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
ENV = dict(os.environ, IQS_BENCH_CPU_QUBITS="14", OMP_NUM_THREADS="2")


def json_lines(text):
    out = []
    for line in text.splitlines():
        line = line.strip()
        if line.startswith("{") and line.endswith("}"):
            out.append(json.loads(line))
    return out


def check_line(rec, steps, warmup, gpus):
    assert rec["impl"] == "reference" and rec["metric"] == "effective_GBps" and rec["unit"] == "GB/s"
    assert rec["steps"] == steps and rec["warmup"] == warmup and rec["n_gpus"] == gpus
    assert rec["higher_is_better"] is True and rec["dtype"] == "f64" and rec["vs_baseline"] is None
    assert rec["value"] > 0 and rec["ms_per_step"] > 0
    assert rec["config"]["workload"] and rec["config"]["qubits_sampled"] == 14
    cb = rec["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == rec["value"] and cb["sample"]
    assert rec["e2e"] == {"value": rec["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert set(rec["paths"]) == {"generic", "specialize_v2"}  # the reference at its best: both paths timed
    assert rec["gpu_launches"] == 0


def test_reference_arm_single_process():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "3"],
                       capture_output=True, text=True, timeout=600, env=ENV, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    recs = json_lines(r.stdout)
    assert len(recs) == 1
    check_line(recs[0], 2, 3, 1)


def test_reference_arm_under_torchrun_only_rank0_prints():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29731",
           os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "2", "--warmup", "3"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=ENV, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    recs = json_lines(r.stdout)
    assert len(recs) == 1
    check_line(recs[0], 2, 3, 2)


def test_own_arm_fails_loudly_without_a_gpu():
    """no silent CPU path behind bench.py: without a CUDA device the run ends with an error and prints no line"""
    import pytest
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "3"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode != 0 and "no CUDA device" in r.stderr
    assert json_lines(r.stdout) == []
