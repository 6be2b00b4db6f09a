mkdir -p gpurun_out
python -m pytest tests/test_capi_gpu.py tests/test_dropin_gpu.py -m gpu -q -x 2>&1 | tail -3
for d in build/variants/*/; do name=$(basename $d); for dbg in io noio; do echo "=== $name $dbg"; IQS_B200_FUSED_DEBUG=$dbg IQS_B200_LIB=$PWD/$d/libiqs_b200.so timeout 300 python tools/kbench.py --n 32 --reps 3 --ops fused 2>&1 | grep -E "fused1 |fused32|bench_layer|fused12_gen|fused12_x|fused12_real"; done; done 2>&1 | tee gpurun_out/r02v_params_vs_smem_n32.log
