#!/bin/bash
# GPU side of tools/fused_variants.sh: time every variant with kbench (fused ops only).
cd "$(dirname "$0")/.."
N=${1:-30}
mkdir -p gpurun_out
for d in build/variants/*/; do
  name=$(basename $d)
  echo "=== $name"
  IQS_B200_LIB=$PWD/$d/libiqs_b200.so timeout 300 python tools/kbench.py --n $N --reps 3 --ops fused --out gpurun_out/fusedvar_$name.json 2>&1 | grep -E "fused|bench_layer"
done
