"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/iqs_ref_driver, built from
/root/reference by oracle/Makefile).  Run in the build container:  python tests/golden/make_golden.py
Single-threaded so that the OpenMP reductions (ComputeNorm) are order-deterministic."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from pkg import circuits as C  # noqa: E402
from progs import random_program  # noqa: E402
import __graft_entry__ as g  # noqa: E402

orc = g.load_oracle()
assert orc.have_ref_driver(), "build oracle/_ref first (python __graft_entry__.py)"


def save(name, prog, psi):
    r = orc.run_reference(prog, state=psi, threads=1)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), n=prog.n, ops=np.frombuffer(prog.ops.tobytes(), dtype=np.uint8),
                        state_in=psi, state_out=r["state"], scalars=r["scalars"], map=r["map"])
    print(name, prog.n, len(prog), "ops")


# every op kind on a tiny register
p = random_program(5, 250, 101)
for q in range(5):
    p.prob(q)
p.expect([0, 2, 4], [1, 2, 3]).expect1(1, 1).expect1(3, 2).norm()
save("ref_allkinds_5q", p, C.random_state(5, 101))

# permuted register + measurement
p = random_program(8, 120, 102)
p.permute([3, 0, 7, 1, 6, 2, 5, 4])
p.extend(random_program(8, 120, 103))
p.emuswap(1, 6)
p.extend(random_program(8, 40, 104))
for q in range(8):
    p.prob(q)
p.collapse(2, 1).normalize().norm()
save("ref_permuted_8q", p, C.random_state(8, 102))

# fusion on
p = C.Program(9).mode(C.FUSION_ON, 5)
p.extend(random_program(9, 200, 105, kinds="basic"))
p.mode(C.FUSION_OFF)
save("ref_fusion_9q", p, C.random_state(9, 105))

# the benchmark circuit family, QFT and Heisenberg
save("ref_layered_10q", C.layered_random(10, 3), C.random_state(10, 106))
save("ref_qft_10q", C.qft(10), C.random_state(10, 777))
base1 = np.zeros(256, dtype=np.complex128)
base1[1] = 1
save("ref_heisenberg_8q", C.heisenberg_step(8), base1)
