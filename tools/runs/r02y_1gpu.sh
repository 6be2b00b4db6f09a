mkdir -p gpurun_out
for d in build/variants/*/; do name=$(basename $d); echo "=== $name"; IQS_B200_LIB=$PWD/$d/libiqs_b200.so python -m pytest tests/test_capi_gpu.py -m gpu -q -x -k "fused" 2>&1 | tail -1
IQS_B200_LIB=$PWD/$d/libiqs_b200.so timeout 300 python tools/kbench.py --n 32 --reps 3 --ops fused 2>&1 | grep -E "fused1 |fused32|bench_layer|fused12_gen|fused12_x"; done 2>&1 | tee gpurun_out/r02y_variants_n32.log
