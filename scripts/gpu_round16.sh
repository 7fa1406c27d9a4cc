#!/bin/bash
mkdir -p gpurun_out
echo "== old"
(cd _old && timeout 300 python scripts/gemm_bench.py conv320 conv640 tconv320 2>&1 | tail -n 3)
echo "== new"
timeout 600 python scripts/gemm_bench.py 2>&1 | tail -n 20
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x 2>&1 | tail -n 4
