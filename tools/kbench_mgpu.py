"""NVLink-side kernel timing: gates on global qubits through the C ABI, one process per GPU.

launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/kbench_mgpu.py [--local-qubits 30]
Reports, per op, ms (max over ranks), algorithmic NVLink GB/s per GPU per direction and
the fraction of the measured 770 GB/s peer-copy reference / 900 GB/s nominal.
"""
import argparse
import json
import math
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

pkg = g.load_package()
capi, C = pkg.capi, pkg.circuits


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--local-qubits", dest="m", type=int, default=30, help="local qubits per GPU")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    box = [capi.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    ctx = capi.Context(rank, world, box[0], device=lr)
    M = a.m
    L = 1 << M
    k = int(math.log2(world))
    st = ctx.alloc(L, tmp_amps=L // 4)
    st.share()
    st.fill_random(7, rank * L)
    X = np.array([0, 0, 1, 0, 1, 0, 0, 0.0])
    rows = []

    def timeit(name, fn, bytes_dir):
        fn()
        ctx.barrier()
        dist.barrier()
        best = 1e30
        for _ in range(a.reps):
            ctx.timer_start()
            fn()
            ms = ctx.timer_stop()
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            best = min(best, float(t.item()))
        gbs = bytes_dir / (best * 1e-3) / 1e9
        if rank == 0:
            rows.append({"op": name, "ms": best, "nvlink_gbs_per_dir": gbs, "frac_770": gbs / 770.0, "frac_900": gbs / 900.0})
            print(f"{name:28s} {best:9.3f} ms  {gbs:7.1f} GB/s per GPU per direction  {100 * gbs / 770:5.1f}% of 770 measured  {100 * gbs / 900:5.1f}% of 900 nominal", flush=True)

    for gq in range(k):
        timeit(f"gate1_global(bit {gq})", lambda gq=gq: st.gate1_global(M, M + gq, C.G_FIXED), 16.0 * L)
    timeit("cgate1_global(c=5)", lambda: st.cgate1_global(M, 5, M, X), 8.0 * L)
    timeit("cgate1_global(c=M-1)", lambda: st.cgate1_global(M, M - 1, M, X), 8.0 * L)
    timeit("cgate1_global(c=0)", lambda: st.cgate1_global(M, 0, M, X), 8.0 * L)
    timeit("swap_global(local 3, glob)", lambda: st.swap2x2_global(M, 3, M, X), 8.0 * L)
    timeit("swap_global(local M-1, glob)", lambda: st.swap2x2_global(M, M - 1, M + k - 1, X), 8.0 * L)
    if k >= 2:
        timeit("swap_global(glob, glob)", lambda: st.swap2x2_global(M, M, M + 1, X), 16.0 * L)
    # the placement layer's primitive: k local positions <-> k global positions in one pass;
    # each call is undone by the next one (same arguments), so the state is preserved
    for kk in range(1, min(k, 3) + 1):
        gl = [M + j for j in range(kk)]
        for name, lo in (("high", [M - 1 - j for j in range(kk)]), ("mid", [12 + 3 * j for j in range(kk)]), ("low5", [5 + j for j in range(kk)])):
            timeit(f"exchange_bits(k={kk}, {name} {lo})", lambda lo=lo, gl=gl: st.exchange_bits(M, lo, gl), 16.0 * L * (1.0 - 2.0 ** -kk))
    timeit("permute_global(pair swap)", lambda: st.permute_global(rank ^ 1, rank ^ 1), 16.0 * L)
    timeit("permute_global(shift 1)", lambda: st.permute_global((rank + 1) % world, (rank - 1) % world), 16.0 * L)
    timeit("local gate1(pos 10) [ref]", lambda: st.gate1(10, C.G_FIXED), 32.0 * L)
    if rank == 0 and a.out:
        os.makedirs(os.path.dirname(a.out), exist_ok=True)
        json.dump({"world": world, "M": M, "rows": rows}, open(a.out, "w"), indent=1)
    st.free()
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
