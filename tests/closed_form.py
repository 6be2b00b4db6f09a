"""Closed-form circuit for parity checks at sizes no CPU reference can hold (test helper).

|0...0> -> one seeded random unitary u_q per qubit (a product state: the amplitude of index i is
prod_q u_q[bit_q(i), 0]) -> CNOT(q, q+1) for q = 0..n-2 in that order (basis state i moves to its
prefix-XOR j, so amp'(j) = amp(i) with bit_q(i) = bit_q(j) ^ bit_{q-1}(j)) -> controlled phase
gates CPhase(a, b, theta) (amp'(j) *= e^{i theta} when bits a and b of j are set).

Every qubit carries a different random unitary, so an index error at ANY bit position -- in
particular at bits 28..34, which no oracle comparison reaches -- changes the sampled amplitudes by
O(|amp|) and cannot cancel (unlike G.G^dagger round trips).  Style of the reference's exact-position
checks, unit_test/include/qureg_permute_test.hpp:60-259.
"""
import math

import numpy as np

from progs import random_unitary


class ClosedForm:
    def __init__(self, n, seed=2026, ncphase=8):
        self.n = n
        rng = np.random.Generator(np.random.MT19937(seed))
        self.u = [random_unitary(rng) for _ in range(n)]
        self.cphase = []
        for _ in range(ncphase):
            a, b = (int(x) for x in rng.permutation(n)[:2])
            self.cphase.append((a, b, float(rng.uniform(0, 2 * math.pi))))
        # always exercise the top bits
        if n >= 3:
            self.cphase.append((n - 1, 0, 0.7))
            self.cphase.append((1, n - 2, 1.9))

    def gates(self):
        """[(kind, control, target, m8)] with kind 0 = 1-qubit gate, 1 = controlled gate."""
        X = np.array([0, 0, 1, 0, 1, 0, 0, 0.0])
        out = []
        for q in range(self.n):
            out.append((0, 0, q, np.ascontiguousarray(self.u[q]).ravel().view(np.float64).copy()))
        for q in range(self.n - 1):
            out.append((1, q, q + 1, X))
        for a, b, th in self.cphase:
            out.append((1, a, b, np.array([1, 0, 0, 0, 0, 0, math.cos(th), math.sin(th)])))
        return out

    def program(self, C, samples, fused=False):
        p = C.Program(self.n)
        if fused:
            p.mode(C.FUSION_ON, 10)
        for kind, c, t, m in self.gates():
            if kind == 0:
                p.gate1(t, m)
            else:
                p.cgate1(c, t, m)
        if fused:
            p.mode(C.FUSION_OFF)
        for j in samples:
            p.get_amp(int(j))
        return p

    def samples(self, count, seed=5):
        """Random indices plus indices that force every high bit (and runs of high bits)."""
        n = self.n
        rng = np.random.Generator(np.random.MT19937(seed))
        idx = [int(x) for x in rng.integers(0, 1 << n, size=count, dtype=np.uint64)]
        for b in range(n):
            idx.append(1 << b)
            idx.append(((1 << n) - 1) ^ (1 << b))
            idx.append(int(rng.integers(0, 1 << n, dtype=np.uint64)) | (1 << b) | (1 << (n - 1)))
        idx += [0, (1 << n) - 1]
        return idx

    def amplitude(self, j):
        n = self.n
        a = 1.0 + 0.0j
        prev = 0
        for q in range(n):
            bj = (j >> q) & 1
            a *= self.u[q][bj ^ prev, 0]
            prev = bj
        for c, t, th in self.cphase:
            if (j >> c) & 1 and (j >> t) & 1:
                a *= complex(math.cos(th), math.sin(th))
        return a

    def check(self, samples, got, tol=1e-12):
        want = np.array([self.amplitude(j) for j in samples])
        got = np.asarray(got)
        err = np.abs(got - want)
        scale = np.max(np.abs(want))
        worst = int(np.argmax(err))
        assert err[worst] <= tol, f"index {samples[worst]:#x}: got {got[worst]}, closed form {want[worst]} (|d| = {err[worst]:.3e})"
        # 1e-12 absolute is loose for amplitudes of magnitude 2^-(n/2): also require 1e-9 of the largest sample
        assert err[worst] <= 1e-9 * scale + 1e-300, f"relative error {err[worst] / scale:.3e} at index {samples[worst]:#x}"
        return float(err[worst]), float(scale)
