mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x --durations=5 2>&1 | tee gpurun_out/r02g_pytest_gpu.log | tail -15
python bench.py --steps 5 --warmup 3 > gpurun_out/r02g_bench_n1.json 2> gpurun_out/r02g_bench_n1.err; tail -3 gpurun_out/r02g_bench_n1.err; cat gpurun_out/r02g_bench_n1.json
ncu --set full --clock-control none --import-source on -k regex:k_pairs_w2 -s 3 -c 2 -o gpurun_out/r02_kpairs_n32 python tools/kbench.py --n 32 --reps 1 --ops gate1 > gpurun_out/r02g_ncu.log 2>&1; tail -2 gpurun_out/r02g_ncu.log
python tools/ncu_traffic.py gpurun_out/r02_kpairs_n32.ncu-rep --kernel k_pairs_w2 --qubits 32 --out gpurun_out/r02_traffic.json
ls -la gpurun_out/r02_kpairs_n32.ncu-rep
