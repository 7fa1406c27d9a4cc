#!/bin/bash
mkdir -p gpurun_out
for t in 0 0x050000 0x0a0000; do
  echo "== T2V_GEMM_TUNE=$t"
  T2V_GEMM_TUNE=$t timeout 200 python scripts/gemm_bench.py lin320_res lin320 qkv320 geglu320 lin640_res geglu640 qkv640 q640 conv320 tconv320 2>&1 | tail -n 10
done | tee gpurun_out/r02_pair_sweep.txt
