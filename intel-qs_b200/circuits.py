"""Circuit ("program") descriptions shared by the tests, bench.py and the drivers.

A program is a numpy structured array of ``OP_DTYPE`` records, byte-compatible with
``struct iqs_op`` in oracle/iqs_program.h; every record names PROGRAM qubits exactly like the
public ``iqs::QubitRegister`` methods (reference include/qureg.hpp:212-331).

The generators restate the reference's own workloads:
  * ``qft``               examples/quantum_fourier_transform.cpp:29-74
  * ``layered_random``    BASELINE.json configs[1] (SURVEY.md 8d "Config 2"): one random 1-qubit
                          gate per qubit drawn from {G, H, RX, RY, RZ, sqrtX, sqrtY, T}, then CNOTs
                          on alternating even/odd neighbour pairs
  * ``heisenberg_step``   examples/heisenberg_dynamics.cpp:81-128
"""
import math

import numpy as np

OP_DTYPE = np.dtype(
    [("kind", "<i4"), ("q0", "<i4"), ("q1", "<i4"), ("q2", "<i4"), ("p", "<f8", (40,))], align=False
)
assert OP_DTYPE.itemsize == 336

HEADER_DTYPE = np.dtype(
    [("magic", "<u4"), ("num_qubits", "<u4"), ("nops", "<u4"), ("init", "<u4"), ("base_index", "<u8"), ("reserved", "<u8")]
)
MAGIC = 0x50535149

# op kinds (oracle/iqs_program.h)
GATE1, CGATE1, SWAPLIKE, DIAG, GATE2, TOFFOLI = 1, 2, 3, 4, 5, 6
H, X, Y, Z, SQRTX, SQRTY, SQRTZ, T, RX, RY, RZ, RXY = 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21
CH, CX, CY, CZ, CSQRTZ, CRX, CRY, CRZ, CPHASE = 30, 31, 32, 33, 34, 35, 36, 37, 38
SWAP, ISWAP, SQRTISWAP, FOURTHROOTISWAP = 40, 41, 42, 43
PROB, EXPECT, NORM, NORMALIZE, COLLAPSE, EXPECT1, ENTROPY, GOOGLESTATS, GETAMP = 50, 51, 52, 53, 54, 55, 56, 57, 58
PERMUTE, EMUSWAP = 60, 61
FUSION_ON, FUSION_OFF, SPEC_ON, SPEC_OFF, SPEC2_ON, SPEC2_OFF = 70, 71, 72, 73, 74, 75

# The fixed "random" unitary used throughout the reference (benchmarks/basic_code_for_scaling.cpp:100-104,
# unit_test/include/apply_1q_gate_test.hpp:241-245, tutorials/get_started_with_IQS.cpp:208-211).
G_FIXED = np.array(
    [0.592056606032915, 0.459533060553574, -0.314948020757856, -0.582328159830658,
     0.658235557641767, 0.070882241549507, 0.649564427121402, 0.373855203932477]
)


class Program:
    """Append-only list of ops."""

    def __init__(self, num_qubits):
        self.n = int(num_qubits)
        self._ops = []

    def _add(self, kind, q0=0, q1=0, q2=0, p=()):
        rec = np.zeros((), dtype=OP_DTYPE)
        rec["kind"], rec["q0"], rec["q1"], rec["q2"] = kind, q0, q1, q2
        p = np.asarray(p, dtype=np.float64).ravel()
        rec["p"][: p.size] = p
        self._ops.append(rec)
        return self

    # generic
    def gate1(self, q, m):
        return self._add(GATE1, q, p=_m8(m))

    def cgate1(self, c, t, m):
        return self._add(CGATE1, c, t, p=_m8(m))

    def swaplike(self, q1, q2, m):
        return self._add(SWAPLIKE, q1, q2, p=_m8(m))

    def diag(self, q1, q2, d4):
        d = np.asarray(d4, dtype=np.complex128).ravel()
        assert d.size == 4
        return self._add(DIAG, q1, q2, p=d.view(np.float64))

    def gate2(self, qh, ql, m4):
        m = np.asarray(m4, dtype=np.complex128).reshape(16)
        return self._add(GATE2, qh, ql, p=m.view(np.float64))

    def toffoli(self, c1, c2, t):
        return self._add(TOFFOLI, c1, c2, t)

    def named1(self, kind, q, *angles):
        return self._add(kind, q, p=angles)

    def named2(self, kind, c, t, *angles):
        return self._add(kind, c, t, p=angles)

    def prob(self, q):
        return self._add(PROB, q)

    def expect(self, qubits, observables):
        p = np.zeros(40)
        p[: len(qubits)] = qubits
        p[16 : 16 + len(observables)] = observables
        return self._add(EXPECT, len(qubits), p=p)

    def expect1(self, q, obs):
        return self._add(EXPECT1, q, obs)

    def norm(self):
        return self._add(NORM)

    def entropy(self):
        return self._add(ENTROPY)

    def google_stats(self):
        return self._add(GOOGLESTATS)

    def get_amp(self, global_index):
        assert 0 <= int(global_index) < (1 << 53)
        return self._add(GETAMP, p=[float(int(global_index))])

    def normalize(self):
        return self._add(NORMALIZE)

    def collapse(self, q, value):
        return self._add(COLLAPSE, q, int(bool(value)))

    def permute(self, new_map):
        assert len(new_map) == self.n
        return self._add(PERMUTE, p=np.asarray(new_map, dtype=np.float64))

    def emuswap(self, q1, q2):
        return self._add(EMUSWAP, q1, q2)

    def mode(self, kind, arg=0):
        return self._add(kind, arg)

    def extend(self, other):
        self._ops.extend(other._ops)
        return self

    @property
    def ops(self):
        if not self._ops:
            return np.zeros(0, dtype=OP_DTYPE)
        return np.array(self._ops, dtype=OP_DTYPE)

    def __len__(self):
        return len(self._ops)

    def count_gates(self):
        """Number of gate ops (value-returning and mode ops excluded); Toffoli counts as 5."""
        k = self.ops["kind"]
        return int(np.sum(k < 50) + 4 * np.sum(k == TOFFOLI))

    def write(self, path, init=1, base_index=0):
        hdr = np.zeros((), dtype=HEADER_DTYPE)
        hdr["magic"], hdr["num_qubits"], hdr["nops"], hdr["init"], hdr["base_index"] = MAGIC, self.n, len(self), init, base_index
        with open(path, "wb") as f:
            f.write(hdr.tobytes())
            f.write(self.ops.tobytes())


def _m8(m):
    m = np.asarray(m)
    if m.dtype.kind == "c":
        m = np.ascontiguousarray(m, dtype=np.complex128).ravel().view(np.float64)
    m = np.asarray(m, dtype=np.float64).ravel()
    assert m.size == 8
    return m


# ---------------------------------------------------------------------------------------
# workloads
# ---------------------------------------------------------------------------------------
def qft(n):
    """examples/quantum_fourier_transform.cpp:29-74: for i = n-1..0: controlled phase shifts
    (control j, target i) with angle pi/2^(j-i) for j = n-1..i+1, then H(i); finally swaps."""
    prog = Program(n)
    for i in range(n - 1, -1, -1):
        for j in range(n - 1, i, -1):
            k = j - i
            # TM2x2 phaseshift: (1,0),(0,0),(0,0), (cos(pi/2^k), sin(pi/2^k))   (:57-62)
            ang = math.pi / float(1 << k)
            m = np.array([1, 0, 0, 0, 0, 0, math.cos(ang), math.sin(ang)], dtype=np.float64)
            prog.cgate1(j, i, m)
        prog.named1(H, i)
    for i in range(n // 2):
        prog.named2(SWAP, i, n - 1 - i)
    return prog


def layered_random(n, layers, seed=20971):
    """Config 2: per layer one random 1-qubit gate on every qubit followed by CNOT(q, q+1) on
    alternating even / odd pairs.  numpy's MT19937 seeded as stated; angles ~ U[0, 2pi)."""
    rng = np.random.Generator(np.random.MT19937(seed))
    prog = Program(n)
    kinds = ["G", H, RX, RY, RZ, SQRTX, SQRTY, T]
    for layer in range(layers):
        for q in range(n):
            k = kinds[int(rng.integers(0, len(kinds)))]
            if k == "G":
                prog.gate1(q, G_FIXED)
            elif k in (RX, RY, RZ):
                prog.named1(k, q, float(rng.uniform(0.0, 2.0 * math.pi)))
            else:
                prog.named1(k, q)
        for q in range(layer % 2, n - 1, 2):
            prog.named2(CX, q, q + 1)
    return prog


def heisenberg_step(n, with_expectations=True):
    """examples/heisenberg_dynamics.cpp:81-128 (register starts in "base" index 1): RY(3.14159/6) on
    all qubits, <Z> on every qubit, one Trotter step (per bond i: RX(i, g*J*dt), CX(i,i+1),
    RZ(i+1, J*dt), CX(i,i+1); then RX on the last qubit) with J = g = 1, dt = 0.1, <Z> again."""
    J, g, dt = 1.0, 1.0, 0.1
    prog = Program(n)
    for q in range(n):
        prog.named1(RY, q, 3.14159 / 6.0)
    if with_expectations:
        for q in range(n):
            prog.expect1(q, 3)
    for q in range(n - 1):
        prog.named1(RX, q, g * J * dt)
        prog.named2(CX, q, q + 1)
        prog.named1(RZ, q + 1, J * dt)
        prog.named2(CX, q, q + 1)
    prog.named1(RX, n - 1, g * J * dt)
    if with_expectations:
        for q in range(n):
            prog.expect1(q, 3)
    return prog


def scaling_sweep(n, gates_per_qubit=1):
    """benchmarks/basic_code_for_scaling.cpp:119-136: the fixed G on every qubit in turn."""
    prog = Program(n)
    for q in range(n):
        for _ in range(gates_per_qubit):
            prog.gate1(q, G_FIXED)
    return prog


def random_state(n, seed=777):
    """re, im ~ U[-1, 1) from numpy MT19937(seed), normalised (SURVEY.md 8d Config 1)."""
    rng = np.random.Generator(np.random.MT19937(seed))
    v = rng.uniform(-1.0, 1.0, size=(1 << n, 2))
    s = np.ascontiguousarray(v).view(np.complex128).ravel()
    s /= np.linalg.norm(s)
    return s
