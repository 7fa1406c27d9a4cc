#!/bin/bash
mkdir -p gpurun_out
CASES="lin320_res lin320 geglu320 lin640_res lin1280_res lin1280_l3 conv320 conv640 conv1280 conv1280_l3 tconv320 tconv1280_l3"
for tune in 0 2 3 0x100 0x103; do
  echo "== T2V_GEMM_TUNE=$tune"
  T2V_GEMM_TUNE=$tune timeout 300 python scripts/gemm_bench.py $CASES 2>&1 | tail -n 12
done
