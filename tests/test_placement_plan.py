"""CPU: the planning half of the placement layer (iqsb_plan_placement, csrc/placement.cu) driven by a
Python mirror of the host library's queue loop (intel-qs_b200/src/placement.cpp: Enqueue / RunQueue /
BringLocal).  Checks the policy on the BASELINE workloads: how many multi-bit exchanges a circuit
costs, that a gate only ever runs with its non-diagonal target on a local bit, and that the placement
stays a permutation."""
import numpy as np
import pytest

from pkg import capi, circuits as C

DIAG_KINDS = {C.RZ, C.T, C.Z, C.SQRTZ, C.CZ, C.CSQRTZ, C.CRZ, C.CPHASE}
CTRL_KINDS = {C.CGATE1, C.CH, C.CX, C.CY, C.CZ, C.CSQRTZ, C.CRX, C.CRY, C.CRZ, C.CPHASE}


def gates_of(prog):
    """(kind, control, target, diagonal) per gate op of a program with the identity qubit map."""
    out = []
    for op in prog.ops:
        k = int(op["kind"])
        if k >= 50 or k == C.SWAP:
            continue
        if k == C.GATE1 or k == C.CGATE1:
            p = op["p"]
            diag = p[2] == 0 and p[3] == 0 and p[4] == 0 and p[5] == 0
        else:
            diag = k in DIAG_KINDS
        if k in CTRL_KINDS:
            out.append((1, int(op["q0"]), int(op["q1"]), diag))
        else:
            out.append((0, 0, int(op["q0"]), diag))
    return out


class Sim:
    def __init__(self, n, M, lookahead=128, min_evict=5):
        self.n, self.M, self.lookahead, self.min_evict = n, M, lookahead, min_evict
        self.place = list(range(n))
        self.queue = []
        self.exchanges = []  # (gate index at which it happened, k)
        self.last_use = [0] * n
        self.clock = 0
        self.done = 0

    def blocks(self, g):
        return self.place[g[2]] >= self.M and not g[3]

    def run(self, count):
        i = 0
        while i < count:
            g = self.queue[i]
            if self.blocks(g):
                ev, br = capi.plan_placement(self.place, self.M, self.queue[i : i + 512], 1 << g[2], self.last_use, self.min_evict)
                assert len(ev) == len(br) >= 1 and g[2] in br
                for e, b in zip(ev, br):
                    assert self.place[e] < self.M <= self.place[b] and self.place[e] >= self.min_evict
                    self.place[e], self.place[b] = self.place[b], self.place[e]
                self.exchanges.append((self.done + i, len(ev)))
                assert sorted(self.place) == list(range(self.n))
                assert not self.blocks(g)
                continue
            i += 1
        self.done += count
        del self.queue[:count]

    def feed(self, gates):
        for g in gates:
            self.queue.append(g)
            self.clock += 1
            self.last_use[g[2]] = self.clock
            if g[0] == 1:
                self.last_use[g[1]] = self.clock
            if len(self.queue) >= 2 * self.lookahead:
                self.run(len(self.queue) - self.lookahead)
        self.run(len(self.queue))


def link_bytes(exchanges, M):
    """bytes per rank and direction, in units of 16 B * 2^M"""
    return sum(1.0 - 2.0 ** -k for _, k in exchanges)


def test_layered_circuit_costs_one_exchange_per_layer():
    """BASELINE configs[1] on 8 GPUs (35 qubits, 32 local): per layer of 35 one-qubit gates + CNOTs the
    reference's layout pays ~2.25 dense gates on rank bits (16*L each way) plus the CNOTs that reach
    them; the planner needs about one 3-bit exchange (14*L)."""
    n, M, layers = 35, 32, 12
    sim = Sim(n, M)
    sim.feed(gates_of(C.layered_random(n, layers)))
    per_layer = len(sim.exchanges) / layers
    assert per_layer <= 1.35, sim.exchanges
    # traffic: reference layout = every non-diagonal gate whose target is a rank bit moves 1 (dense) or 1/2 (controlled) unit
    ref = 0.0
    for kind, c, t, d in gates_of(C.layered_random(n, layers)):
        if t >= M and not d:
            ref += 1.0 if kind == 0 or c >= M else 0.5
    ours = link_bytes(sim.exchanges, M) * 0.5 * 2  # exchange of k bits moves (1 - 2^-k) * L amplitudes = that many 16-byte units
    assert ours < 0.5 * ref, (ours, ref)


def test_qft_needs_a_handful_of_exchanges():
    """BASELINE configs[2]: QFT at 34 qubits on 8 GPUs.  Controlled phases are diagonal (free on rank
    bits); only the H gates on the three global qubits need them local, once."""
    n, M = 34, 31
    sim = Sim(n, M)
    sim.feed(gates_of(C.qft(n)))
    assert len(sim.exchanges) <= 2, sim.exchanges


def test_eager_mode_without_lookahead_still_terminates():
    n, M = 12, 9
    sim = Sim(n, M, lookahead=0, min_evict=2)
    rng = np.random.default_rng(1)
    gates = [(int(rng.integers(0, 2)), 0, 0, False) for _ in range(300)]
    gates = []
    for _ in range(300):
        a, b = (int(x) for x in rng.permutation(n)[:2])
        gates.append((int(rng.integers(0, 2)), a, b, bool(rng.integers(0, 4) == 0)))
    sim.lookahead = 1
    sim.feed(gates)
    assert sim.done == len(gates)


def test_protected_positions_are_brought_in_and_kept():
    n, M = 10, 7
    place = list(range(n))
    ev, br = capi.plan_placement(place, M, [], protect_mask=(1 << 8) | (1 << 9) | (1 << 3), min_evict_bit=0)
    assert sorted(br) == [8, 9] and 3 not in ev and all(e < M for e in ev)
