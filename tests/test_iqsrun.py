"""tools/iqsrun, the one-process-per-GPU launcher (the image has no mpirun): environment it hands to the
ranks, exit-code propagation, the deadline.  Host logic only -- the ranks here are plain Python processes."""
import os
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
IQSRUN = os.path.join(os.path.dirname(HERE), "tools", "iqsrun")


def launch(n, code, timeout=None, limit=60):
    cmd = [sys.executable, IQSRUN, "-n", str(n)] + (["--timeout", str(timeout)] if timeout else []) + [sys.executable, "-c", code]
    t0 = time.time()
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=limit)
    return r.returncode, r.stdout, time.time() - t0


def test_every_rank_gets_its_coordinates_and_one_rendezvous_file():
    rc, out, _ = launch(4, "import os; print(os.environ['IQS_RANK'], os.environ['IQS_NRANKS'], os.environ['IQS_LOCAL_RANK'], os.environ['IQS_UID_FILE'])")
    assert rc == 0
    rows = sorted(line.split() for line in out.splitlines())
    assert [r[:3] for r in rows] == [[str(k), "4", str(k)] for k in range(4)]
    files = {r[3] for r in rows}
    assert len(files) == 1  # one file for the whole launch ...
    assert not os.path.exists(files.pop())  # ... that does not outlive it (rank 0 of a real run creates it)


def test_first_failing_rank_decides_the_exit_code_and_the_others_are_stopped():
    code = "import os, sys, time\nif os.environ['IQS_RANK'] == '1': sys.exit(7)\ntime.sleep(30)"
    rc, _, took = launch(3, code)
    assert rc == 7 and took < 15  # the sleeping ranks were killed, not waited for


def test_deadline():
    rc, _, took = launch(2, "import time; time.sleep(30)", timeout=1.5)
    assert rc == 124 and took < 15


def test_usage_errors():
    for args in ([], ["-n", "2"], ["--bogus", "x"]):
        r = subprocess.run([sys.executable, IQSRUN] + args, capture_output=True, text=True, timeout=30)
        assert r.returncode != 0 and "usage" in (r.stderr + r.stdout)
