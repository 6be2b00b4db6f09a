mkdir -p gpurun_out
python -m pytest tests/test_capi_gpu.py tests/test_pybind_gpu.py -m gpu -q -x -k "permute or reference_binding or one_amplitude or marginals or pauli" 2>&1 | tail -3
for b in 1 0; do echo "== bulk=$b"; IQS_B200_PERMUTE_BULK=$b python tools/kbench.py --n 32 --reps 3 --ops permute 2>&1 | grep permute; done 2>&1 | tee gpurun_out/r02o_permute_bulk_n32.log
python tools/kbench.py --n 32 --reps 3 --ops prob,pauli,norm,parity --out gpurun_out/r02o_kbench_reduce_n32.json 2>&1 | tee gpurun_out/r02o_kbench_reduce_n32.log | tail -12
( for f in 0 11; do python tools/run_configs.py heisenberg --n 32 --fusion $f 2>&1 | tail -1; IQS_B200_ONE_SWEEP=0 python tools/run_configs.py heisenberg --n 32 --fusion $f 2>&1 | tail -1; done ) | tee gpurun_out/r02o_configs_heisenberg.log
