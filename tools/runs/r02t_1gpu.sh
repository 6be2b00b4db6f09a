mkdir -p gpurun_out
python -m pytest tests/test_capi_gpu.py -m gpu -q -x -k "fused or large_state" 2>&1 | tail -2
for dbg in io noio; do echo "== $dbg"; IQS_B200_FUSED_DEBUG=$dbg python tools/kbench.py --n 32 --reps 3 --ops fused 2>&1 | grep -E "fused1 |fused32|bench_layer|fused12_gen|fused12_x|fused12_real"; done 2>&1 | tee gpurun_out/r02t_fused_noio_n32.log
