mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python -m pytest tests/test_multigpu.py tests/test_capi_gpu.py -m gpu -q -x -k "qft_sharded or fusion_sharded or fused_gate_classes or reads_between" 2>&1 | tee gpurun_out/r02af_sanity_2gpu.log | tail -4
