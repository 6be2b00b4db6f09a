mkdir -p gpurun_out
python -m pytest tests/test_multigpu.py -m gpu -q -x -rs --durations=5 2>&1 | tee gpurun_out/r02aa_pytest_multigpu_2gpu.log | tail -10
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02aa_bench_n2.json 2> gpurun_out/r02aa_bench_n2.err; python -c "
import json; d=json.load(open('gpurun_out/r02aa_bench_n2.json')); print(d['value'], d['e2e']['value'], d['e2e']['fused']['gates_per_s'], d['roofline_nvlink']['achieved'], d['kernel_classes'])"
