"""Build recipe: nvcc for sm_100a, in-tree outputs (they travel to the GPU box with gpurun).

  csrc/*.cu                -> lib/libiqs_b200.so   (CUDA kernels + C ABI, include/iqsb.h)
  src/*.cpp                -> lib/libiqs.so        (iqs::QubitRegister host API over the C ABI)
  pybind/intelqs_py.cpp    -> lib/intelqs_py*.so   (Python module, same surface as the reference's)
  oracle/driver.cpp        -> bin/iqs_b200_driver  (the oracle's driver source against OUR library)
"""
import concurrent.futures as cf
import glob
import os
import shutil
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OBJ = os.path.join(ROOT, "build")
LIB = os.path.join(HERE, "lib")
BIN = os.path.join(HERE, "bin")

NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
# the image exports CXX=/opt/gcc wrapper; the system compiler is the one with a complete runtime
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-ccbin", CXX, "-Wno-deprecated-gpu-targets",
]
CXX_FLAGS = ["-O2", "-std=c++14", "-fPIC", "-Wall", "-Wno-unused-variable", "-Wno-sign-compare"]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("build step failed:\n  " + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    return r


def build_cuda(verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(LIB, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(HERE, "csrc", "*.cu")))
    hdrs = glob.glob(os.path.join(HERE, "csrc", "*.cuh")) + [os.path.join(ROOT, "include", "iqsb.h")]
    jobs = []
    objs = []
    for s in srcs:
        o = os.path.join(OBJ, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        if _newer(o, [s] + hdrs):
            jobs.append([NVCC] + NVCC_FLAGS + ["-c", s, "-o", o])
    with cf.ThreadPoolExecutor(max_workers=8) as ex:
        list(ex.map(_run, jobs))
    so = os.path.join(LIB, "libiqs_b200.so")
    if jobs or _newer(so, objs):
        _run([NVCC, "-shared", "-ccbin", CXX, "-Wno-deprecated-gpu-targets", "-o", so] + objs + ["-ldl"])
    if verbose:
        print(f"[build] {so} ({len(jobs)} objects recompiled)")
    return so


def build_host(verbose=False):
    """libiqs.so: the re-authored iqs::QubitRegister API (host C++) on top of the C ABI."""
    srcs = sorted(glob.glob(os.path.join(HERE, "src", "*.cpp")))
    if not srcs:
        return None
    hdrs = glob.glob(os.path.join(HERE, "include", "*.hpp")) + [os.path.join(ROOT, "include", "iqsb.h")]
    inc = ["-I" + os.path.join(HERE, "include"), "-I" + os.path.join(ROOT, "include")]
    jobs, objs = [], []
    for s in srcs:
        o = os.path.join(OBJ, "host_" + os.path.basename(s)[:-4] + ".o")
        objs.append(o)
        if _newer(o, [s] + hdrs):
            jobs.append([CXX] + CXX_FLAGS + inc + ["-c", s, "-o", o])
    with cf.ThreadPoolExecutor(max_workers=8) as ex:
        list(ex.map(_run, jobs))
    so = os.path.join(LIB, "libiqs.so")
    if jobs or _newer(so, objs):
        _run([CXX, "-shared", "-o", so] + objs + ["-L" + LIB, "-liqs_b200", "-Wl,-rpath,$ORIGIN"])
    if verbose:
        print(f"[build] {so} ({len(jobs)} objects recompiled)")
    return so


def build_driver(verbose=False):
    src = os.path.join(ROOT, "oracle", "driver.cpp")
    if not os.path.exists(os.path.join(LIB, "libiqs.so")):
        return None
    os.makedirs(BIN, exist_ok=True)
    exe = os.path.join(BIN, "iqs_b200_driver")
    deps = [src, os.path.join(ROOT, "oracle", "iqs_program.h"), os.path.join(LIB, "libiqs.so")] + glob.glob(os.path.join(HERE, "include", "*.hpp"))
    if _newer(exe, deps):
        _run([CXX] + CXX_FLAGS + ["-I" + os.path.join(HERE, "include"), "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "oracle"),
                                  "-o", exe, src, "-L" + LIB, "-liqs", "-liqs_b200", "-Wl,-rpath,$ORIGIN/../lib"])
    if verbose:
        print(f"[build] {exe}")
    # the pool-of-states scenario (tests/pool_check.cpp), launched by tests/test_multigpu.py
    pool_src = os.path.join(ROOT, "tests", "pool_check.cpp")
    pool_exe = os.path.join(BIN, "pool_check")
    if os.path.exists(pool_src) and _newer(pool_exe, [pool_src, os.path.join(LIB, "libiqs.so")] + glob.glob(os.path.join(HERE, "include", "*.hpp"))):
        _run([CXX] + CXX_FLAGS + ["-I" + os.path.join(HERE, "include"), "-I" + os.path.join(ROOT, "include"), "-o", pool_exe, pool_src,
                                  "-L" + LIB, "-liqs", "-liqs_b200", "-Wl,-rpath,$ORIGIN/../lib"])
    return exe


def build_pybind(verbose=False):
    src = os.path.join(HERE, "pybind", "intelqs_py.cpp")
    if not os.path.exists(src) or not os.path.exists(os.path.join(LIB, "libiqs.so")):
        return None
    import pybind11

    ext = sysconfig.get_config_var("EXT_SUFFIX")
    so = os.path.join(LIB, "intelqs_py" + ext)
    deps = [src, os.path.join(LIB, "libiqs.so")] + glob.glob(os.path.join(HERE, "include", "*.hpp"))
    if _newer(so, deps):
        _run([CXX] + CXX_FLAGS + ["-shared", "-fvisibility=hidden", "-I" + pybind11.get_include(), "-I" + sysconfig.get_paths()["include"],
                                  "-I" + os.path.join(HERE, "include"), "-I" + os.path.join(ROOT, "include"), "-o", so, src,
                                  "-L" + LIB, "-liqs", "-liqs_b200", "-Wl,-rpath,$ORIGIN"])
    if verbose:
        print(f"[build] {so}")
    return so


def build_oracle(verbose=False):
    _run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "-j8"])
    if verbose:
        print("[build] oracle/liboracle.so" + (" + oracle/_ref" if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "iqs_ref_driver")) else ""))


def build_all(verbose=False):
    build_cuda(verbose)
    build_host(verbose)
    build_driver(verbose)
    build_pybind(verbose)
    build_oracle(verbose)


if __name__ == "__main__":
    build_all(verbose=True)
    sys.exit(0)
