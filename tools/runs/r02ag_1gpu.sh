mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python -m pytest tests -m gpu -q -x 2>&1 | tee gpurun_out/r02ag_pytest_gpu_1gpu.log | tail -3
