#!/bin/bash
# where do the K = 320 level-0 GEMMs lose their time?  timing experiments with parts of the epilogue disabled (results are wrong)
mkdir -p gpurun_out
for t in 0x0 0x1000 0x2000 0x3000 0x4000 0x7000 0x2 0x3 0x4; do
  echo "== T2V_GEMM_TUNE=$t"; T2V_GEMM_TUNE=$t timeout 120 python scripts/gemm_bench.py lin320_res lin320 qkv320 geglu320 ff2_320 2>&1 | tail -n 5
done
for bn in 128 256; do echo "== BN=$bn"; BN=$bn timeout 120 python scripts/gemm_bench.py lin320_res lin320 qkv320 2>&1 | tail -n 3; done
