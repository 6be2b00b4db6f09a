mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python -m pytest tests -m gpu -q -x -rs --durations=5 2>&1 | tee gpurun_out/r02ab_pytest_gpu_1gpu.log | tail -12
python bench.py > gpurun_out/r02ab_bench_n1.json 2> gpurun_out/r02ab_bench_n1.err; tail -2 gpurun_out/r02ab_bench_n1.err; python -c "
import json; d=json.load(open('gpurun_out/r02ab_bench_n1.json')); print(d['value'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['fused']['gates_per_s'], d['cpu_baseline'])"
( time python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r02ab_bench_reference_n1.json 2> gpurun_out/r02ab_bench_reference_n1.err ) 2>&1 | tail -3; cat gpurun_out/r02ab_bench_reference_n1.json
