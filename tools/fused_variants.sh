#!/bin/bash
# Build tuning variants of the fused kernel (compile-time knobs of csrc/kernels_fused.cu: IQSB_FUSED_TILE,
# IQSB_FUSED_REGBITS; the CTA shape is chosen at run time) as
# build/variants/<name>/libiqs_b200.so; tools/kbench.py picks one with IQS_B200_LIB=<path>.
set -e
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.build()" >/dev/null
variants=(
  "default:"
  "r4:-DIQSB_FUSED_REGBITS=4"
)
for v in "${variants[@]}"; do
  name=${v%%:*}; flags=${v#*:}
  d=build/variants/$name; mkdir -p $d
  echo "== $name $flags"
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -ccbin /usr/bin/g++ -Wno-deprecated-gpu-targets \
    -Xptxas -v $flags -c intel-qs_b200/csrc/kernels_fused.cu -o $d/kernels_fused.o 2>&1 | grep -A2 "k_fusedId" | grep -E "Used|spill" || true
  objs=$(ls build/*.o | grep -v "host_\|kernels_fused.o")
  nvcc -shared -ccbin /usr/bin/g++ -Wno-deprecated-gpu-targets -o $d/libiqs_b200.so $objs $d/kernels_fused.o -ldl
done
