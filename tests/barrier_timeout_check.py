"""Launched on 2 GPUs by tests/test_multigpu.py (tools/iqsrun -n 2): rank 1 stays away from the second
barrier; rank 0 must get IQSB_ERR_PEER after IQS_B200_BARRIER_TIMEOUT_S instead of spinning for ever."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

capi = g.load_package().capi
rank, uidf = int(os.environ["IQS_RANK"]), os.environ["IQS_UID_FILE"]
if rank == 0:
    uid = capi.unique_id()
    with open(uidf + ".tmp", "wb") as f:
        f.write(uid)
    os.rename(uidf + ".tmp", uidf)
else:
    t0 = time.time()
    while not os.path.exists(uidf):
        assert time.time() - t0 < 60
        time.sleep(0.01)
    uid = open(uidf, "rb").read()
os.environ["IQS_B200_BARRIER_TIMEOUT_S"] = "3"
ctx = capi.Context(rank, 2, uid, device=rank)
ctx.barrier()  # everybody is here: no error
ctx.check()
if rank == 1:
    time.sleep(15)
    os._exit(0)
t0 = time.time()
try:
    ctx.barrier()
except capi.IqsbError as e:
    dt = time.time() - t0
    print(f"TIMEOUT_OK after {dt:.1f} s: {e}", flush=True)
    os._exit(0 if 2.0 < dt < 12.0 else 3)
print("the barrier returned without its partner and without an error", flush=True)
os._exit(2)
