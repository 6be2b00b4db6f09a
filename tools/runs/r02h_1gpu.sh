mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x --durations=5 2>&1 | tee gpurun_out/r02h_pytest_gpu.log | tail -8
python tools/kbench.py --n 32 --reps 3 --ops fused --out gpurun_out/r02h_kbench_fused_n32.json 2>&1 | tee gpurun_out/r02h_kbench_fused_n32.log | grep -E "fused1 |fused_hi|layer|fused12"
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r02h_bench_n1.json 2> gpurun_out/r02h_bench_n1.err; python -c "
import json; d=json.load(open('gpurun_out/r02h_bench_n1.json')); print(d['value'], d['e2e']['value'], json.dumps(d['e2e']['fused']))"
