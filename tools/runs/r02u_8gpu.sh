mkdir -p gpurun_out
# measured when a permutation of a sharded register was lazy by default; that scheme is now opt-in
IQS_B200_LAZY_PERMUTE=1 timeout 300 python tools/run_configs.py reorder --n 35 --ranks 8 2>&1 | tail -2 | tee gpurun_out/r02u_config4_8gpu.log
