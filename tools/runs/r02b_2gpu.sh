mkdir -p gpurun_out
python -m pytest tests/test_multigpu.py -m gpu -q -rs -x --durations=8 2>&1 | tee gpurun_out/r02b_pytest_multigpu_2gpu.log | tail -30
python - <<'P' 2>&1 | tail -30
import os, sys, time, subprocess
sys.path.insert(0, ".")
import __graft_entry__ as g
C = g.load_package().circuits
orc = g.load_oracle()
p = C.layered_random(12, 2)
os.environ["IQS_DRIVER_TRACE"] = "1"
os.environ["IQS_B200_TRACE"] = "1"
t0 = time.time()
r = orc.run_driver("intel-qs_b200/bin/iqs_b200_driver", p, init=2, want_state=False, launcher=[sys.executable, "tools/iqsrun", "-n", "2"])
print("wall", time.time() - t0)
P
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/kbench_mgpu.py --local-qubits 30 --out gpurun_out/r02b_kbench_mgpu_n2.json 2>&1 | grep -v Warning | tail -30
for f in 0 11; do for pl in 1 0; do IQS_B200_PLACEMENT=$pl python tools/run_configs.py layered --n 31 --ranks 2 --fusion $f 2>&1 | tail -1; done; done | tee gpurun_out/r02b_layered_2gpu.log
