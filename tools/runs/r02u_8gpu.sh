mkdir -p gpurun_out
timeout 300 python tools/run_configs.py reorder --n 35 --ranks 8 2>&1 | tail -2 | tee gpurun_out/r02u_config4_8gpu.log
