"""CPU: the partition of the multi-bit local<->global exchange (iqsb_plan_exchange, the pure host
function behind iqsb_exchange_bits).  Every rank is emulated with a numpy shard; executing the plans
of all ranks must equal the bit permutation "swap position lpos[j] with position gpos[j]" of the
global vector, every amplitude must be written at most once, and the link traffic must be
(1 - 2^-k) * L amplitudes per rank and direction.  Also run as world_size-2/4 gloo processes."""
import os
import socket
import sys

import numpy as np
import pytest

from pkg import capi

HERE = os.path.dirname(os.path.abspath(__file__))


def swap_bits_global(vec, n, pairs):
    idx = np.arange(1 << n, dtype=np.int64)
    dst = idx.copy()
    for a, b in pairs:
        ba, bb = (dst >> a) & 1, (dst >> b) & 1
        dst = dst & ~((1 << a) | (1 << b)) | (bb << a) | (ba << b)
    out = np.empty_like(vec)
    out[dst] = vec
    return out


def block_indices(M, lpos, pattern, split_bit, split_val):
    k = np.arange(1 << M, dtype=np.int64)
    keep = ((k >> split_bit) & 1) == split_val
    for p in lpos:
        keep &= ((k >> p) & 1) == ((pattern >> p) & 1)
    return k[keep]


def emulate(world, M, lpos, gpos, shards):
    """apply every rank's plan to `shards` (list of arrays) the way k_exchange does."""
    L = 1 << M
    new = [s.copy() for s in shards]
    written = [np.zeros(L, dtype=np.int32) for _ in range(world)]
    for r in range(world):
        pl = capi.plan_exchange(r, world, M, lpos, gpos)
        assert pl.npartners == (1 << len(lpos)) - 1
        assert pl.link_amps == pl.npartners * (L >> len(lpos))
        for p in range(pl.npartners):
            q = pl.partner[p]
            ia = block_indices(M, lpos, pl.mine[p], pl.split_bit, pl.split_val[p])
            ib = block_indices(M, lpos, pl.theirs[p], pl.split_bit, pl.split_val[p])
            assert len(ia) == len(ib) == pl.amps_per_partner
            new[r][ia] = shards[q][ib]
            new[q][ib] = shards[r][ia]
            written[r][ia] += 1
            written[q][ib] += 1
    for w in written:
        assert w.max() <= 1
    return new, written


@pytest.mark.parametrize("world", [2, 4, 8])
def test_exchange_plan_is_the_bit_swap(world):
    g = int(np.log2(world))
    rng = np.random.default_rng(world)
    for M in (4, 6, 9):
        n = M + g
        L = 1 << M
        vec = np.arange(1 << n, dtype=np.float64) + 0.25
        for k in range(1, min(g, 3) + 1):
            for _ in range(6):
                lpos = [int(x) for x in rng.permutation(M)[:k]]
                gpos = [int(x) + M for x in rng.permutation(g)[:k]]
                shards = [vec[r * L : (r + 1) * L] for r in range(world)]
                new, written = emulate(world, M, lpos, gpos, shards)
                want = swap_bits_global(vec, n, list(zip(lpos, gpos)))
                assert np.array_equal(np.concatenate(new), want), (M, lpos, gpos)
                # every amplitude whose exchanged local bits differ from its rank bits moved exactly once
                moved = sum(int(w.sum()) for w in written)
                assert moved == world * ((1 << k) - 1) * (L >> k)
    with pytest.raises(capi.IqsbError):
        capi.plan_exchange(0, world, 4, [1], [2])  # 2 is not a global position
    with pytest.raises(capi.IqsbError):
        capi.plan_exchange(0, world, 1, [0], [1])  # needs k + 1 local qubits


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, M, cases, out_q):
    sys.path.insert(0, HERE)
    import torch
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = M + int(np.log2(world))
    L = 1 << M
    vec = np.arange(1 << n, dtype=np.float64) * 3.0 + 1.0
    shard = vec[rank * L : (rank + 1) * L].copy()
    ok, msg = True, ""
    for lpos, gpos in cases:
        vec = swap_bits_global(vec, n, list(zip(lpos, gpos)))
        gathered = [torch.zeros(L, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(gathered, torch.from_numpy(shard))  # "peer memory"
        peers = [t.numpy() for t in gathered]
        pl = capi.plan_exchange(rank, world, M, lpos, gpos)
        remote_writes = []
        for p in range(pl.npartners):
            q = pl.partner[p]
            ia = block_indices(M, lpos, pl.mine[p], pl.split_bit, pl.split_val[p])
            ib = block_indices(M, lpos, pl.theirs[p], pl.split_bit, pl.split_val[p])
            remote_writes.append((q, ib, shard[ia].copy()))
            shard[ia] = peers[q][ib]
        allw = [None] * world
        dist.all_gather_object(allw, remote_writes)
        for w in allw:
            for owner, idx, vals in w:
                if owner == rank:
                    shard[idx] = vals
        if not np.array_equal(shard, vec[rank * L : (rank + 1) * L]):
            ok, msg = False, f"exchange {lpos}<->{gpos}: rank {rank} differs"
            break
    flags = [None] * world
    dist.all_gather_object(flags, (ok, msg))
    if rank == 0:
        out_q.put(flags)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_exchange_over_gloo(world):
    import torch.multiprocessing as mp

    M = 6
    g = int(np.log2(world))
    cases = [([3], [M]), ([0], [M + g - 1]), ([M - 1], [M])]
    if world == 4:
        cases += [([2, 5], [M + 1, M]), ([0, 1], [M, M + 1]), ([4], [M + 1])]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, M, cases, q)) for r in range(world)]
    for p in procs:
        p.start()
    flags = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
    for ok, msg in flags:
        assert ok, msg
