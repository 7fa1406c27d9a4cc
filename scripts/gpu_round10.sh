#!/bin/bash
mkdir -p gpurun_out
for tune in 0 0x800; do
  echo "== T2V_GEMM_TUNE=$tune"
  T2V_GEMM_TUNE=$tune timeout 300 python scripts/gemm_bench.py vae512 big1280 conv320 2>&1 | tail -n 3
done

timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "pairs" 2>&1 | tail -n 3
T2V_GEMM_TUNE=0 timeout 300 python scripts/gemm_bench.py lin320_res qkv320 geglu320 ff2_320 lin640_res geglu640 conv640 tconv320 2>&1 | tail -n 8
