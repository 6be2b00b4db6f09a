mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02m_bench_n8.json 2> gpurun_out/r02m_bench_n8.err; tail -3 gpurun_out/r02m_bench_n8.err; cat gpurun_out/r02m_bench_n8.json
timeout 500 python -m pytest tests/test_multigpu.py -m gpu -q -x -rs --durations=8 -k "placement-8 or (reads_between_moves and 8) or (fusion_sharded and 8) or (closed_form and 35) or (qft_sharded and 8)" 2>&1 | tee gpurun_out/r02m_pytest_multigpu_8gpu.log | tail -14
( for r in 8 4 2; do for f in 0 11; do timeout 300 python tools/run_configs.py qft --n 34 --ranks $r --fusion $f 2>&1 | tail -1; done; done
  for f in 0 11; do timeout 300 python tools/run_configs.py reorder --n 35 --ranks 8 --fusion $f 2>&1 | tail -2; done ) | tee gpurun_out/r02m_configs_8gpu.log
