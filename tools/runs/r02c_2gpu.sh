mkdir -p gpurun_out
python -m pytest tests/test_multigpu.py -m gpu -q -rs --durations=5 2>&1 | tee gpurun_out/r02c_pytest_multigpu_2gpu.log | tail -40
for f in 0 11; do for pl in 1 0; do IQS_B200_PLACEMENT=$pl python tools/run_configs.py layered --n 31 --ranks 2 --fusion $f 2>&1 | tail -1; done; done | tee gpurun_out/r02c_layered_2gpu.log
