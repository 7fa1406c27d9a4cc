#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "pairs" > gpurun_out/t_pairs.log 2>&1; echo "pair tests rc=$?"; tail -n 12 gpurun_out/t_pairs.log | cut -c1-400
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x > gpurun_out/t_kernels.log 2>&1; echo "kernel tests rc=$?"; tail -n 4 gpurun_out/t_kernels.log | cut -c1-400
CASES="lin320_res lin320 qkv320 geglu320 ff2_320 lin640_res geglu640 lin1280_res conv320 conv640 conv1280 tconv320"
for tune in 0 0x800; do
  echo "== T2V_GEMM_TUNE=$tune"
  T2V_GEMM_TUNE=$tune timeout 300 python scripts/gemm_bench.py $CASES 2>&1 | tail -n 12
done
