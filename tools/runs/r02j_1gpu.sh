mkdir -p gpurun_out
for n in 30 32; do
ncu --set full --clock-control none -k regex:k_fused -s 1 -c 1 -o gpurun_out/r02j_kfused1_n$n -f python tools/kbench.py --n $n --reps 1 --ops fused > gpurun_out/r02j_ncu_$n.log 2>&1; tail -1 gpurun_out/r02j_ncu_$n.log
done
nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,clocks_event_reasons.sw_power_cap --format=csv -lms 100 > gpurun_out/r02j_clocks.csv &
SMI=$!
python tools/kbench.py --n 32 --reps 6 --ops fused 2>&1 | grep -E "fused1 |fused32"
kill $SMI
sort gpurun_out/r02j_clocks.csv | uniq -c | sort -rn | head -12
