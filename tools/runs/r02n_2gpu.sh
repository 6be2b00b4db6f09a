mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -rs --durations=6 2>&1 | tee gpurun_out/r02n_pytest_gpu_2gpu.log | tail -16
