"""CPU: the tile-phase planner of PermuteLocalQubits (iqsb_plan_permute, pure host code).
Every phase is emulated with numpy exactly as the kernel runs it (tile-local scatter through
"shared memory"), and the composition must equal the requested bit permutation."""
import numpy as np
import pytest

from pkg import capi


def apply_phase(vec, nbits, pos, dstslot):
    """in-place tile phase: inside every tile, slot t moves to slot sigma(t)"""
    idx = np.arange(1 << nbits, dtype=np.int64)
    dst = idx.copy()
    for k, p in enumerate(pos):
        dst &= ~(1 << p)
    for k, p in enumerate(pos):
        dst |= ((idx >> p) & 1) << pos[dstslot[k]]
    out = np.empty_like(vec)
    out[dst] = vec
    return out


def expected(vec, nbits, dst_bit):
    idx = np.arange(1 << nbits, dtype=np.int64)
    j = np.zeros_like(idx)
    for b in range(nbits):
        j |= ((idx >> b) & 1) << dst_bit[b]
    out = np.empty_like(vec)
    out[j] = vec
    return out


def check(dst_bit):
    n = len(dst_bit)
    phases = capi.plan_permute(dst_bit)
    vec = np.arange(1 << n, dtype=np.int64)
    got = vec
    for pos, dstslot in phases:
        assert pos == sorted(pos) and len(set(pos)) == len(pos) and len(pos) <= 12
        assert sorted(dstslot) == list(range(len(pos)))  # a permutation of the tile bits
        assert pos[: min(4, n)] == list(range(min(4, n)))  # low bits always in the tile: coalesced runs
        got = apply_phase(got, n, pos, dstslot)
    assert np.array_equal(got, expected(vec, n, dst_bit))
    return len(phases)


@pytest.mark.parametrize("n", [1, 2, 3, 5, 8, 12, 13, 16, 18])
def test_random_permutations(n):
    rng = np.random.default_rng(n)
    for _ in range(12):
        check([int(x) for x in rng.permutation(n)])


def test_structured_permutations():
    for n in (14, 17, 20):
        assert check(list(range(n))) == 0  # identity: nothing to do
        nrev = check(list(range(n))[::-1])  # full reversal
        nrot = check([(b + 1) % n for b in range(n)])  # one long cycle
        swp = list(range(n))
        swp[1], swp[n - 2] = swp[n - 2], swp[1]
        assert check(swp) == 1
        assert nrev <= (n // 2 + 3) // 4 + 1 and nrot <= (n + 6) // 7 + 1


def test_phase_counts_at_benchmark_sizes():
    """32 local qubits: a full reversal needs 4 in-place passes, a single swap 1 (planner only)."""
    n = 32
    assert len(capi.plan_permute(list(range(n))[::-1])) <= 4
    swp = list(range(n))
    swp[0], swp[31] = swp[31], swp[0]
    assert len(capi.plan_permute(swp)) == 1
    # A phase holds 12 positions, 4 of them always the lowest: it can put at most 8 positions >= 4 into their
    # final place, so ceil(moved / 8) phases are needed whatever the schedule.  The planner is within one of that.
    rng = np.random.default_rng(0)
    for _ in range(200):
        perm = [int(x) for x in rng.permutation(n)]
        moved = sum(1 for b in range(4, n) if perm[b] != b)
        assert -(-moved // 8) <= len(capi.plan_permute(perm)) <= -(-moved // 8) + 1
