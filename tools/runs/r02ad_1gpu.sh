mkdir -p gpurun_out
python -m pytest tests/test_capi_gpu.py tests/test_dropin_gpu.py tests/test_fullsize_gpu.py tests/test_pybind_gpu.py tests/test_examples_dropin_gpu.py -m gpu -q -x 2>&1 | tail -3
timeout 300 python tools/kbench.py --n 32 --reps 3 --ops fused --out gpurun_out/r02ad_kbench_fused_n32.json 2>&1 | grep -E "fused1 |fused32|bench_layer|fused12_|fused_layer"
for f in 11; do python tools/run_configs.py heisenberg --n 32 --fusion $f 2>&1 | tail -1; python tools/run_configs.py qft --n 32 --fusion $f 2>&1 | tail -1; done | tee gpurun_out/r02ad_configs_1gpu.log
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r02ad_bench_n1.json 2> gpurun_out/r02ad_bench_n1.err; python -c "
import json; d=json.load(open('gpurun_out/r02ad_bench_n1.json')); print(d['value'], d['e2e']['value'], d['e2e']['fused']['gates_per_s'], d['e2e']['fused_fma']['gates_per_s'])"
