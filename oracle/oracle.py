"""ctypes face of the CPU oracle (oracle/iqs_oracle.c) and of the reference driver.

TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product never imports this module.
"""
import ctypes
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
REF_DRIVER = os.path.join(HERE, "_ref", "iqs_ref_driver")

_lib = None

_u64 = ctypes.c_uint64
_dp = ctypes.POINTER(ctypes.c_double)


def build():
    """Compile liboracle.so (and oracle/_ref when /root/reference is present)."""
    subprocess.run(["make", "-s", "-C", HERE, "-j8"], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = ctypes.CDLL(LIB_PATH)
        L.oracle_run.restype = ctypes.c_int
        L.oracle_run.argtypes = [ctypes.c_uint, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        L.oracle_prob1.restype = ctypes.c_double
        L.oracle_parity_expect.restype = ctypes.c_double
        L.oracle_norm2.restype = ctypes.c_double
        L.oracle_maxabsdiff.restype = ctypes.c_double
        L.oracle_l2diff.restype = ctypes.c_double
        _lib = L
    return _lib


def _state(a):
    a = np.ascontiguousarray(a, dtype=np.complex128)
    return a, a.ctypes.data_as(ctypes.c_void_p)


def _m(m):
    m = np.asarray(m)
    if m.dtype.kind == "c":
        m = np.ascontiguousarray(m, dtype=np.complex128).ravel().view(np.float64)
    m = np.ascontiguousarray(m, dtype=np.float64).ravel()
    return m, m.ctypes.data_as(ctypes.c_void_p)


# --- kernel-level primitives (positions, in place on a complex128 numpy array) -----------
def gate1(state, pos, m, sind=0, eind=None):
    s, p = _state(state)
    mm, mp = _m(m)
    lib().oracle_gate1(p, _u64(sind), _u64(len(s) if eind is None else eind), ctypes.c_uint(pos), mp)
    return s


def cgate1(state, cpos, tpos, m, sind=0, eind=None):
    s, p = _state(state)
    mm, mp = _m(m)
    lib().oracle_cgate1(p, _u64(sind), _u64(len(s) if eind is None else eind), ctypes.c_uint(cpos), ctypes.c_uint(tpos), mp)
    return s


def swap2x2(state, pos1, pos2, m):
    s, p = _state(state)
    mm, mp = _m(m)
    lib().oracle_swap2x2(p, _u64(len(s)), ctypes.c_uint(pos1), ctypes.c_uint(pos2), mp)
    return s


def diag2(state, pos1, pos2, d):
    s, p = _state(state)
    dd, dp = _m(d)
    lib().oracle_diag2(p, _u64(len(s)), ctypes.c_uint(pos1), ctypes.c_uint(pos2), dp)
    return s


def gate2(state, pos_high, pos_low, m16):
    s, p = _state(state)
    mm = np.ascontiguousarray(np.asarray(m16, dtype=np.complex128).reshape(16)).view(np.float64)
    lib().oracle_gate2(p, _u64(len(s)), ctypes.c_uint(pos_high), ctypes.c_uint(pos_low), mm.ctypes.data_as(ctypes.c_void_p))
    return s


def scale(state, f, start=0, end=None):
    s, p = _state(state)
    ff = np.array([complex(f).real, complex(f).imag])
    lib().oracle_scale(p, _u64(start), _u64(len(s) if end is None else end), ff.ctypes.data_as(ctypes.c_void_p))
    return s


def prob1(state, pos):
    s, p = _state(state)
    return lib().oracle_prob1(p, _u64(len(s)), ctypes.c_uint(pos))


def parity_expect(state, mask, glb_start=0):
    s, p = _state(state)
    return lib().oracle_parity_expect(p, _u64(len(s)), _u64(mask), _u64(glb_start))


def norm2(state):
    s, p = _state(state)
    return lib().oracle_norm2(p, _u64(len(s)))


def overlap(state, psi):
    s, p = _state(state)
    t, q = _state(psi)
    out = np.zeros(2)
    lib().oracle_overlap(p, q, _u64(len(s)), out.ctypes.data_as(ctypes.c_void_p))
    return complex(out[0], out[1])


def maxabsdiff(a, b, f=1.0):
    s, p = _state(a)
    t, q = _state(b)
    ff = np.array([complex(f).real, complex(f).imag])
    return lib().oracle_maxabsdiff(p, q, _u64(len(s)), ff.ctypes.data_as(ctypes.c_void_p))


def l2diff(a, b):
    s, p = _state(a)
    t, q = _state(b)
    return lib().oracle_l2diff(p, q, _u64(len(s)))


def collapse(state, pos, value):
    s, p = _state(state)
    lib().oracle_collapse(p, _u64(len(s)), ctypes.c_uint(pos), ctypes.c_int(int(bool(value))))
    return s


def any_above(state, pos, tol):
    s, p = _state(state)
    out = (ctypes.c_int * 2)()
    lib().oracle_any_above(p, _u64(len(s)), ctypes.c_uint(pos), ctypes.c_double(tol), out)
    return int(out[0]), int(out[1])


def axpy(a, b, f=1.0):
    s, p = _state(a)
    t, q = _state(b)
    ff = np.array([complex(f).real, complex(f).imag])
    lib().oracle_axpy(p, q, _u64(len(s)), ff.ctypes.data_as(ctypes.c_void_p))
    return s


# --- program level ----------------------------------------------------------------------
def run_program(n, state, ops):
    """Replay `ops` (circuits.OP_DTYPE records) on a copy of `state`.
    Returns (final_state in data order, scalars, qubit->position map)."""
    s = np.array(state, dtype=np.complex128, copy=True)
    assert s.size == 1 << n
    ops = np.ascontiguousarray(ops)
    cap = 16 + 11 * len(ops)
    scal = np.zeros(cap)
    qmap = np.zeros(n, dtype=np.uint64)
    ns = lib().oracle_run(n, s.ctypes.data_as(ctypes.c_void_p), ops.ctypes.data_as(ctypes.c_void_p), len(ops),
                          scal.ctypes.data_as(ctypes.c_void_p), cap, qmap.ctypes.data_as(ctypes.c_void_p))
    if ns < 0:
        raise ValueError("oracle: unknown op in program")
    return s, scal[:ns].copy(), qmap.astype(np.int64)


def have_ref_driver():
    return os.path.exists(REF_DRIVER)


def run_driver(driver, program, state=None, init=1, base_index=0, threads=None, repeat=1, want_state=True, extra_env=None, launcher=None):
    """Run a driver binary (reference or drop-in) on `program`.
    Returns dict(state, scalars, map, seconds)."""
    n = program.n
    env = dict(os.environ)
    if threads:
        env["OMP_NUM_THREADS"] = str(threads)
    if extra_env:
        env.update(extra_env)
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as td:
        pf = os.path.join(td, "prog.bin")
        if state is not None:
            init = 0
            np.ascontiguousarray(state, dtype=np.complex128).tofile(os.path.join(td, "in.bin"))
        program.write(pf, init=init, base_index=base_index)
        cmd = list(launcher or []) + [driver, pf, "--scalars-out", os.path.join(td, "scal.bin"), "--map-out", os.path.join(td, "map.bin"), "--repeat", str(repeat)]
        if state is not None:
            cmd += ["--state-in", os.path.join(td, "in.bin")]
        if want_state:
            cmd += ["--state-out", os.path.join(td, "out.bin")]
        r = subprocess.run(cmd, env=env, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"driver failed ({r.returncode}): {r.stdout[-2000:]}\n{r.stderr[-2000:]}")
        secs = settled = None
        for line in r.stdout.splitlines():
            if line.startswith("TIME "):
                f = line.split()
                secs = float(f[1])
                if "SETTLED" in f:
                    settled = float(f[f.index("SETTLED") + 1])
        if env.get("IQS_DRIVER_TRACE"):
            print(r.stderr)
        out = {"seconds": secs, "seconds_settled": settled, "stdout": r.stdout}
        out["scalars"] = np.fromfile(os.path.join(td, "scal.bin"), dtype=np.float64) if os.path.exists(os.path.join(td, "scal.bin")) else np.zeros(0)
        out["map"] = np.fromfile(os.path.join(td, "map.bin"), dtype=np.uint64).astype(np.int64)
        out["state"] = np.fromfile(os.path.join(td, "out.bin"), dtype=np.complex128) if want_state else None
        return out


def run_reference(program, state=None, **kw):
    return run_driver(REF_DRIVER, program, state=state, **kw)
