mkdir -p gpurun_out
python -m pytest tests/test_capi_gpu.py tests/test_dropin_gpu.py tests/test_fullsize_gpu.py -m gpu -q -x 2>&1 | tail -3
for dbg in io noio; do echo "=== $dbg"; IQS_B200_FUSED_DEBUG=$dbg timeout 300 python tools/kbench.py --n 32 --reps 3 --ops fused 2>&1 | grep -E "fused1 |fused32|bench_layer|fused12_gen|fused12_x|fused12_real|fused_layer"; done 2>&1 | tee gpurun_out/r02x_trailing_n32.log
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r02x_bench_n1.json 2> gpurun_out/r02x_bench_n1.err; python -c "
import json; d=json.load(open('gpurun_out/r02x_bench_n1.json')); print(d['value'], d['e2e']['value'], d['e2e']['fused']['gates_per_s'], d['e2e']['fused_fma']['gates_per_s'])"
