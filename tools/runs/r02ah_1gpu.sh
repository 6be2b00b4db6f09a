mkdir -p gpurun_out
python -m pytest tests/test_capi_gpu.py tests/test_dropin_gpu.py tests/test_fullsize_gpu.py tests/test_examples_dropin_gpu.py tests/test_pybind_gpu.py -m gpu -q -x 2>&1 | tail -2
timeout 200 python tools/kbench.py --n 32 --reps 3 --ops fused 2>&1 | grep -E "fused1 |bench_layer|fused12_diag|fused12_gen|fused12_x"
python tools/run_configs.py heisenberg --n 32 --fusion 11 2>&1 | tail -1 | cut -c1-170; python tools/run_configs.py qft --n 32 --fusion 11 2>&1 | tail -1 | cut -c1-150
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r02ah_bench_n1.json 2> gpurun_out/r02ah_bench_n1.err; python -c "
import json; d=json.load(open('gpurun_out/r02ah_bench_n1.json')); print(d['value'], d['e2e']['value'], d['e2e']['fused']['gates_per_s'])"
