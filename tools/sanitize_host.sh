#!/bin/sh
# sanitize_host.sh -- the HOST code of the engine under AddressSanitizer + UBSan, no GPU needed.
#   1. csrc/*.cu is rebuilt into /tmp/iqsb_asan/libiqs_b200.so with the host compiler instrumented; the planner
#      tests (fusion schedule + CPU model of the kernel, permutation phases, exchange partition, placement
#      policy) and a fuzz batch run against it through IQS_B200_LIB;
#   2. src/*.cpp is compiled, instrumented, into the register-free part of the reference's unit tests
#      (tests/reference_host_suite.cpp; needs oracle/_ref/dropin, i.e. a build with /root/reference present).
# Any "runtime error" / "AddressSanitizer" line is a finding.  Round 2: none.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=/tmp/iqsb_asan
CXX=/usr/bin/g++
mkdir -p $OUT/obj
cd $ROOT/intel-qs_b200/csrc
ls *.cu | xargs -P 8 -I{} sh -c "nvcc -gencode arch=compute_100a,code=sm_100a -O1 -g -std=c++17 -Xcompiler -fPIC,-fsanitize=address,-fsanitize=undefined,-fno-omit-frame-pointer -ccbin $CXX -Wno-deprecated-gpu-targets -diag-suppress 550,1886 -I../../include -c {} -o $OUT/obj/\$(basename {} .cu).o"
nvcc -shared -ccbin $CXX -Wno-deprecated-gpu-targets -Xcompiler -fsanitize=address,-fsanitize=undefined -o $OUT/libiqs_b200.so $OUT/obj/*.o -ldl
cd $ROOT
export LD_PRELOAD="$($CXX -print-file-name=libasan.so) $($CXX -print-file-name=libubsan.so)"
export ASAN_OPTIONS=detect_leaks=0:protect_shadow_gap=0
IQS_B200_LIB=$OUT/libiqs_b200.so python -m pytest tests/test_fused_plan.py tests/test_permute_plan.py tests/test_exchange_plan.py tests/test_placement_plan.py -q -s 2>&1 | grep -E "runtime error|AddressSanitizer|passed|failed" | sort | uniq -c
(cd /tmp && IQS_B200_LIB=$OUT/libiqs_b200.so python $ROOT/tools/fuzz_fused_model.py 300000 300 2>&1 | grep -E "runtime error|AddressSanitizer|done|MISMATCH" | sort | uniq -c)
unset LD_PRELOAD
if [ -d $ROOT/oracle/_ref/dropin/unit_test ]; then
  $CXX -O1 -g -std=c++14 -w -fsanitize=address,undefined -fno-sanitize-recover=undefined -DIQS_WITH_NOISE tests/reference_host_suite.cpp intel-qs_b200/src/*.cpp \
    -Itests/gtest_shim -Ioracle/_ref/dropin -Iintel-qs_b200/include -Iinclude -Lintel-qs_b200/lib -liqs_b200 -Wl,-rpath,$ROOT/intel-qs_b200/lib -o $OUT/host_suite
  (cd /tmp && $OUT/host_suite 2>&1 | grep -E "runtime error|ERROR|SUMMARY|PASSED|FAILED")
fi
