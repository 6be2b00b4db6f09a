mkdir -p gpurun_out
python -m pytest tests/test_multigpu.py -m gpu -q -x -rs --durations=5 2>&1 | tee gpurun_out/r02l_pytest_multigpu_2gpu.log | tail -12
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02l_bench_n2.json 2> gpurun_out/r02l_bench_n2.err; tail -5 gpurun_out/r02l_bench_n2.err; cat gpurun_out/r02l_bench_n2.json
