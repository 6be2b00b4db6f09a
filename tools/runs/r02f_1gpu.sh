mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x --durations=8 2>&1 | tee gpurun_out/r02f_pytest_gpu.log | tail -25
bash tools/fused_variants_run.sh 30 2>&1 | tee gpurun_out/r02f_variants.log | grep -E "===|fused1 |fused32|fused12_gen|fused12_x|fused48_gen|fused48_x|layer"
