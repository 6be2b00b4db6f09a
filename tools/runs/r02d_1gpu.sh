mkdir -p gpurun_out
python -m pytest tests/test_capi_gpu.py -m gpu -q -x -k "fused or small_fused or large_state" 2>&1 | tee gpurun_out/r02d_pytest_fused.log | tail -15
python tools/kbench.py --n 30 --reps 3 --ops fused --out gpurun_out/r02d_kbench_fused_n30.json 2>&1 | tee gpurun_out/r02d_kbench_fused_n30.log | tail -40
