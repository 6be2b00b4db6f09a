mkdir -p gpurun_out
python -m pytest tests/test_capi_gpu.py -m gpu -q -x -k "fused or small_fused or large_state" 2>&1 | tee gpurun_out/r02e_pytest_fused.log | tail -5
IQS_B200_LIB=$PWD/build/variants/r4mb2/libiqs_b200.so python -m pytest tests/test_capi_gpu.py -m gpu -q -x -k "fused or small_fused or large_state" 2>&1 | tail -3
bash tools/fused_variants_run.sh 30 2>&1 | tee gpurun_out/r02e_variants.log | grep -E "===|fused12|fused48_gen|fused48_x|layer"
ncu --set full --clock-control none --import-source on -k regex:k_fused -c 8 -o gpurun_out/r02e_kfused_layer_n28 python tools/kbench.py --n 28 --reps 1 --ops fusedprof > gpurun_out/r02e_ncu.log 2>&1; tail -3 gpurun_out/r02e_ncu.log
