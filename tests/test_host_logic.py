"""CPU: host-only classes of the drop-in API (Permutation, TinyMatrix, bit helpers, GateCounter,
RandomNumberGenerator, Environment statics).  tests/host_logic.cpp is compiled against
intel-qs_b200/include + libiqs.so and its output compared with the output of the SAME source
compiled against the reference (live when /root/reference is present, else the committed fixture
tests/golden/host_logic_expected.txt that the live run produced)."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
SRC = os.path.join(HERE, "host_logic.cpp")
EXPECTED = os.path.join(HERE, "golden", "host_logic_expected.txt")


def build_and_run(tmp_path, name, flags):
    exe = str(tmp_path / name)
    r = subprocess.run([CXX, "-O1", "-std=c++14", "-w", SRC, "-o", exe] + flags, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return r.stdout


def ours(tmp_path):
    lib = os.path.join(ROOT, "intel-qs_b200", "lib")
    assert os.path.exists(os.path.join(lib, "libiqs.so")), "run __graft_entry__.build() first"
    return build_and_run(tmp_path, "host_ours", ["-I" + os.path.join(ROOT, "intel-qs_b200", "include"), "-I" + os.path.join(ROOT, "include"),
                                                  "-L" + lib, "-liqs", "-liqs_b200", "-Wl,-rpath," + lib])


def test_host_classes_match_committed_reference_output(tmp_path):
    assert os.path.exists(EXPECTED)
    assert ours(tmp_path) == open(EXPECTED).read()


def test_host_classes_match_live_reference(tmp_path):
    ref_lib = os.path.join(ROOT, "oracle", "_ref")
    if not (os.path.exists("/root/reference/include/qureg.hpp") and os.path.exists(os.path.join(ref_lib, "libiqs_ref.so"))):
        pytest.skip("reference tree not present on this machine")
    ref = build_and_run(tmp_path, "host_ref", ["-fopenmp", "-DUSE_MM_MALLOC", "-I/root/reference/include", "-L" + ref_lib, "-liqs_ref", "-Wl,-rpath," + ref_lib])
    assert ours(tmp_path) == ref
    if os.environ.get("IQS_UPDATE_GOLDEN"):
        open(EXPECTED, "w").write(ref)
