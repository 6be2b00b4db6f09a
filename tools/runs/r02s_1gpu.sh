mkdir -p gpurun_out
for dbg in io noio; do echo "== $dbg"; IQS_B200_FUSED_DEBUG=$dbg python tools/kbench.py --n 32 --reps 3 --ops fused 2>&1 | grep -E "fused1 |fused32|bench_layer|fused12_|fused48_x"; done 2>&1 | tee gpurun_out/r02s_fused_noio_n32.log
