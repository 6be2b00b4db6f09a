mkdir -p gpurun_out
python -m pytest tests/test_capi_gpu.py tests/test_dropin_gpu.py tests/test_fullsize_gpu.py -m gpu -q -x 2>&1 | tail -3
python tools/kbench.py --n 30 --reps 3 --ops fused --out gpurun_out/r02k_kbench_fused_n30.json 2>&1 | grep -E "fused1 |fused_hi1|fused32|bench_layer|fused12_|fused48_x"
for dyn in 1 0; do echo "== n=32 dynamic=$dyn"; IQS_B200_FUSED_DYNAMIC=$dyn python tools/kbench.py --n 32 --reps 3 --ops fused --out gpurun_out/r02k_kbench_fused_n32_dyn$dyn.json 2>&1 | grep -E "fused1 |fused_hi1|fused32|bench_layer|fused12_x|fused12_gen"; done 2>&1 | tee gpurun_out/r02k_dynamic_tiles_n32.log
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r02k_bench_n1.json 2> gpurun_out/r02k_bench_n1.err; python -c "
import json; d=json.load(open('gpurun_out/r02k_bench_n1.json')); print(d['value'], d['e2e']['value'], json.dumps(d['e2e']['fused']), json.dumps(d['e2e']['fused_fma']))"
