mkdir -p gpurun_out
python -m pytest tests/test_capi_gpu.py tests/test_dropin_gpu.py tests/test_fullsize_gpu.py -m gpu -q -x 2>&1 | tail -2
for ph in 1 0; do echo "== phase=$ph"; export IQS_B200_FUSED_PHASE=$ph
timeout 300 python tools/kbench.py --n 32 --reps 3 --ops fused 2>&1 | grep -E "bench_layer|fused12_diag"
python tools/run_configs.py heisenberg --n 32 --fusion 11 2>&1 | tail -1 | cut -c1-170; python tools/run_configs.py qft --n 32 --fusion 11 2>&1 | tail -1 | cut -c1-150; done 2>&1 | tee gpurun_out/r02ae_phase_ab.log
