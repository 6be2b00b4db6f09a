mkdir -p gpurun_out
ncu --set full --clock-control none -k regex:k_fused -s 4 -c 4 -f -o gpurun_out/r02ac_kfused_layer_n28 python tools/kbench.py --n 28 --reps 1 --ops fusedprof > gpurun_out/r02ac_ncu.log 2>&1; tail -1 gpurun_out/r02ac_ncu.log; ls -la gpurun_out/r02ac_kfused_layer_n28.ncu-rep
