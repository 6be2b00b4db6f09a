"""CPU: iqs::ChiMatrix (intel-qs_b200/include/chi_matrix.hpp) is host code -- the reference's
known-answer test for the eigensystem, its container tests and reconstruction of random Hermitian
chi matrices run here without a GPU (tests/chi_matrix_check.cpp)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def test_chi_matrix_host_checks(tmp_path):
    exe = str(tmp_path / "chi_matrix_check")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([cxx, "-O2", "-std=c++14", "-Wall", "-I" + os.path.join(ROOT, "intel-qs_b200", "include"),
                    os.path.join(HERE, "chi_matrix_check.cpp"), "-o", exe], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ALL OK" in r.stdout and "FAILED" not in r.stdout
    assert "OK known_answer" in r.stdout and "OK reconstruct<16>" in r.stdout
