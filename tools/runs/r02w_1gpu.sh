mkdir -p gpurun_out
IQS_B200_FUSED_DEBUG=noio ncu --set full --clock-control none --import-source on -k regex:k_fused -s 1 -c 1 -f -o gpurun_out/r02w_kfused_x_noio_n28 python tools/kbench.py --n 28 --reps 1 --ops fusedx > gpurun_out/r02w_ncu.log 2>&1; tail -1 gpurun_out/r02w_ncu.log
