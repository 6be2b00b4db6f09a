mkdir -p gpurun_out
python -m pytest tests/test_capi_gpu.py -m gpu -q -x -k "fused or large_state" 2>&1 | tail -3
for dyn in 1 0; do echo "== dynamic=$dyn"; IQS_B200_FUSED_DYNAMIC=$dyn python tools/kbench.py --n 32 --reps 3 --ops fused 2>&1 | grep -E "fused1 |fused_hi1|fused32|bench_layer|fused12_x|fused12_gen"; done 2>&1 | tee gpurun_out/r02i_dynamic_tiles_n32.log
