"""CPU, world_size 2 and 4 over gloo: the rank partition of gates on global qubits.

iqsb_plan_global (pure host code, the function the CUDA launchers use) says which pairs each rank
updates in its own and in its partner's shard.  Here every rank is a process holding a numpy shard;
"peer memory" is emulated with gloo all_gather.  The union of the ranks' updates must reproduce the
single-rank oracle exactly, every pair must be owned by exactly one rank, and the bytes a rank moves
over the link must match SURVEY.md 8d."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pkg import capi, circuits as C
from progs import random_unitary

HERE = os.path.dirname(os.path.abspath(__file__))


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def plan_indices(pl, M):
    """local indices k selected by the plan."""
    k = np.arange(1 << M, dtype=np.int64)
    keep = np.ones(k.shape, dtype=bool)
    for i in range(pl.nfix):
        keep &= ((k >> pl.pos[i]) & 1) == pl.val[i]
    return k[keep]


def worker(rank, world, port, n, cases, out_q):
    sys.path.insert(0, HERE)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import __graft_entry__ as g

    orc = g.load_oracle()
    M = n - int(np.log2(world))
    L = 1 << M
    full = C.random_state(n, seed=99)
    want = full.copy()
    shard = full[rank * L : (rank + 1) * L].copy()
    ok = True
    msg = ""
    for kind, p1, p2, m in cases:
        m8 = np.ascontiguousarray(m).ravel().view(np.float64)
        # oracle on the global vector
        if kind == 0:
            orc.gate1(want, p2, m8)
        elif kind == 1:
            orc.cgate1(want, p1, p2, m8)
        else:
            orc.swap2x2(want, p1, p2, m8)
        # "peer memory": everybody can see every shard
        gathered = [torch.zeros(L, dtype=torch.complex128) for _ in range(world)]
        dist.all_gather(gathered, torch.from_numpy(shard))
        shards = [t.numpy().copy() for t in gathered]
        pl = capi.plan_global(kind, rank, world, M, p1, p2)
        writes = []  # (owner rank, local index array, values)
        if pl.active:
            k = plan_indices(pl, M)
            assert len(k) == pl.npairs
            zero_side, one_side = (rank, pl.partner) if pl.role == 0 else (pl.partner, rank)
            i0, i1 = k + pl.extra0, k + pl.extra1
            a0, a1 = shards[zero_side][i0], shards[one_side][i1]
            mm = np.asarray(m, dtype=np.complex128).reshape(2, 2)
            # exact reference order: m00*in0 + m01*in1 via the oracle's own arithmetic on a 2-amp vector
            pair = np.empty(2 * len(k), dtype=np.complex128)
            pair[0::2], pair[1::2] = a0, a1
            orc.gate1(pair, 0, m8)
            writes = [(zero_side, i0, pair[0::2].copy()), (one_side, i1, pair[1::2].copy())]
            del mm
        allw = [None] * world
        dist.all_gather_object(allw, writes)
        touched = np.zeros(L, dtype=np.int32)
        for w in allw:
            for owner, idx, vals in w:
                if owner == rank:
                    shard[idx] = vals
                    touched[idx] += 1
        if touched.max() > 1:
            ok, msg = False, f"kind {kind} ({p1},{p2}): an amplitude was written twice"
        # link traffic: what this rank reads from / writes to the partner, per direction
        if pl.active and pl.link_amps not in (L, L // 2):
            ok, msg = False, "unexpected link byte count"
        if not np.array_equal(shard, want[rank * L : (rank + 1) * L]):
            ok, msg = False, f"kind {kind} ({p1},{p2}): shard differs from the oracle by {np.max(np.abs(shard - want[rank * L:(rank + 1) * L]))}"
        if not ok:
            break
    flags = [None] * world
    dist.all_gather_object(flags, (ok, msg))
    if rank == 0:
        out_q.put(flags)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_global_gate_partition_matches_oracle(world):
    n = 8
    M = n - int(np.log2(world))
    rng = np.random.Generator(np.random.MT19937(5))
    X = np.array([[0, 1], [1, 0]], dtype=np.complex128)
    sym = random_unitary(rng)
    sym[0, 1] = sym[1, 0]
    cases = []
    for g in range(M, n):
        cases.append((0, 0, g, random_unitary(rng)))
        for c in (0, 1, M - 1):
            cases.append((1, c, g, random_unitary(rng)))
        for l in (0, 2, M - 1):
            cases.append((2, l, g, X))
            cases.append((2, l, g, sym))
    if world == 4:
        cases.append((2, M, M + 1, X))
        cases.append((2, M, M + 1, sym))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=worker, args=(r, world, port, n, cases, q)) for r in range(world)]
    for p in procs:
        p.start()
    flags = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
    for ok, msg in flags:
        assert ok, msg


def test_plan_properties_single_process():
    """Every pair of a global gate is owned by exactly one rank; idle ranks own none."""
    for world in (2, 4, 8):
        for M in (1, 2, 5):
            n = M + int(np.log2(world))
            for g in range(M, n):
                owned = 0
                for r in range(world):
                    pl = capi.plan_global(0, r, world, M, 0, g)
                    assert pl.partner == r ^ (1 << (g - M))
                    owned += pl.npairs
                assert owned == (1 << M) * world // 2  # N/2 pairs in total
                if M >= 1:
                    for c in range(M):
                        owned = sum(capi.plan_global(1, r, world, M, c, g).npairs for r in range(world))
                        assert owned == (1 << M) * world // 4
                        owned = sum(capi.plan_global(2, r, world, M, c, g).npairs for r in range(world))
                        assert owned == (1 << M) * world // 4
            if world >= 4:
                owned = sum(capi.plan_global(2, r, world, M, M, M + 1).npairs for r in range(world))
                assert owned == (1 << M) * world // 4
                idle = [r for r in range(world) if not capi.plan_global(2, r, world, M, M, M + 1).active]
                assert all(((r >> 0) & 1) == ((r >> 1) & 1) for r in idle)
    with pytest.raises(capi.IqsbError):
        capi.plan_global(0, 0, 2, 5, 0, 3)  # not a global position


# ---- whole-shard moves when the rank bits are permuted (iqsb_permute_global_bits) ----------------
def permute_worker(rank, world, port, M, perms, out_q):
    sys.path.insert(0, HERE)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    L = 1 << M
    k = int(np.log2(world))
    ok, msg = True, ""
    for dst_bits in perms:
        # shard of rank r: the global indices it holds (canonical placement)
        shard = np.arange(rank * L, (rank + 1) * L, dtype=np.int64)
        src, dst, pairwise, identity = capi.plan_permute_global_bits(rank, world, dst_bits)
        plans = [None] * world
        dist.all_gather_object(plans, (src, dst, pairwise, identity))
        # every rank reached the same decision about the path without talking to anybody
        if len({p[2] for p in plans}) != 1 or len({p[3] for p in plans}) != 1:
            ok, msg = False, f"{dst_bits}: ranks disagree on the path"
        if plans[src][1] != rank or plans[dst][0] != rank:
            ok, msg = False, f"{dst_bits}: source / destination are not inverse of each other"
        if pairwise != all(p[0] == p[1] for p in plans):
            ok, msg = False, f"{dst_bits}: 'pairwise' does not mean every rank trades with one partner"
        gathered = [torch.zeros(L, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(gathered, torch.from_numpy(shard))
        new_shard = gathered[src].numpy()  # "peer memory": pull the source's shard
        # oracle: old global index i lands on j = i with its rank bits permuted
        old = np.arange(world * L, dtype=np.int64)
        r_old = old >> M
        r_new = np.zeros_like(r_old)
        for b in range(k):
            r_new |= ((r_old >> b) & 1) << dst_bits[b]
        j = (r_new << M) | (old & (L - 1))
        want = np.empty_like(old)
        want[j] = old
        if not np.array_equal(new_shard, want[rank * L : (rank + 1) * L]):
            ok, msg = False, f"{dst_bits}: rank {rank} holds the wrong shard"
        if identity != (list(dst_bits) == list(range(k))):
            ok, msg = False, f"{dst_bits}: identity flag"
        if not ok:
            break
    flags = [None] * world
    dist.all_gather_object(flags, (ok, msg))
    if rank == 0:
        out_q.put(flags)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_rank_bit_permutation_plan(world):
    import itertools

    k = int(np.log2(world))
    perms = [list(p) for p in itertools.permutations(range(k))]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=permute_worker, args=(r, world, port, 3, perms, q)) for r in range(world)]
    for p in procs:
        p.start()
    flags = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
    for ok, msg in flags:
        assert ok, msg
