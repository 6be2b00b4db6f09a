mkdir -p gpurun_out
ncu --set full --clock-control none -k regex:k_fused -s 3 -c 3 -f -o gpurun_out/r02p_kfused_layer_n28 python tools/kbench.py --n 28 --reps 1 --ops fusedprof > gpurun_out/r02p_ncu.log 2>&1; tail -2 gpurun_out/r02p_ncu.log
ls -la gpurun_out/r02p_kfused_layer_n28.ncu-rep
python -m pytest tests/test_capi_gpu.py -m gpu -q -x -k "bulk_copy" 2>&1 | tail -2
