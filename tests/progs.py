"""Seeded random programs covering every op kind (test helper)."""
import math

import numpy as np

from pkg import circuits as C


def random_unitary(rng):
    a = rng.normal(size=(2, 2)) + 1j * rng.normal(size=(2, 2))
    q, r = np.linalg.qr(a)
    return q * (np.diag(r) / np.abs(np.diag(r)))


def random_program(n, nops, seed, kinds="all", toffoli=True):
    rng = np.random.Generator(np.random.MT19937(seed))
    p = C.Program(n)
    named1 = [C.H, C.X, C.Y, C.Z, C.SQRTX, C.SQRTY, C.SQRTZ, C.T]
    rot1 = [C.RX, C.RY, C.RZ]
    named2 = [C.CH, C.CX, C.CY, C.CZ, C.CSQRTZ]
    rot2 = [C.CRX, C.CRY, C.CRZ, C.CPHASE]
    swaps = [C.SWAP, C.ISWAP, C.SQRTISWAP, C.FOURTHROOTISWAP]
    for _ in range(nops):
        r = rng.integers(0, 12 if kinds == "all" else 6)
        q = [int(x) for x in rng.permutation(n)[:3]]
        if r == 0:
            p.gate1(q[0], random_unitary(rng))
        elif r == 1:
            p.named1(named1[rng.integers(0, len(named1))], q[0])
        elif r == 2:
            p.named1(rot1[rng.integers(0, 3)], q[0], float(rng.uniform(0, 2 * math.pi)))
        elif r == 3:
            p.cgate1(q[0], q[1], random_unitary(rng))
        elif r == 4:
            p.named2(named2[rng.integers(0, len(named2))], q[0], q[1])
        elif r == 5:
            p.named2(rot2[rng.integers(0, 4)], q[0], q[1], float(rng.uniform(0, 2 * math.pi)))
        elif r == 6:
            p.named2(swaps[rng.integers(0, 4)], q[0], q[1])
        elif r == 7:
            p.diag(q[0], q[1], np.exp(1j * rng.uniform(0, 2 * math.pi, size=4)))
        elif r == 8:
            p.named1(C.RXY, q[0], float(rng.uniform(0, 2 * math.pi)), float(rng.uniform(0, 2 * math.pi)))
        elif r == 9 and toffoli and n >= 3:
            p.toffoli(q[0], q[1], q[2])
        elif r == 10:
            m = random_unitary(rng)
            m[0, 1] = m[1, 0]  # ApplyISwapRotation requires m01 == m10 (reference qureg_applyswap.cpp:58)
            p.swaplike(q[0], q[1], m)
        else:
            p.named1(C.H, q[0])
    return p
