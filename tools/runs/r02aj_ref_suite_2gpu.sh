# the reference's unit-test suite (oracle/_ref/dropin/bin/suite_of_tests) on 2 GPUs, one log per rank.
# The suite reads psi[i] on one rank only in places; permutations settle before they return, and
# IQS_B200_PLACEMENT=0 makes gates on global qubits complete inside the call as well.
# usage: gpurun --gpus 2 --timeout 120 -- 'bash tools/runs/r02aj_ref_suite_2gpu.sh [0|1]'   (argument: IQS_B200_PLACEMENT)
P=${1:-1}
cd oracle/_ref/dropin/bin
IQS_B200_PLACEMENT=$P python /root/repo/tools/iqsrun -n 2 --timeout 60 bash -c \
  "exec ./suite_of_tests > /root/repo/gpurun_out/ref_suite_dropin_2gpu_placement$P.rank\$IQS_RANK.log 2>&1"
echo rc=$?
for r in 0 1; do grep -E "^\[  (PASSED|FAILED)|Failure" /root/repo/gpurun_out/ref_suite_dropin_2gpu_placement$P.rank$r.log | head; done
