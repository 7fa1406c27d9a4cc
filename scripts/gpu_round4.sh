#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/gemm_bench.py > gpurun_out/gemm_bench.log 2>&1; cat gpurun_out/gemm_bench.log
for c in lin320_res geglu320 conv320; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 3 -c 1 -o gpurun_out/prof_$c python scripts/gemm_bench.py $c > gpurun_out/ncu_$c.log 2>&1; echo "ncu $c rc=$?"
done
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "temporal or groupnorm or small_cin" 2>&1 | tail -3
timeout 300 python scripts/shape_profile.py > gpurun_out/shape_profile2.log 2>&1; grep -E "==|groupnorm|attn_short|small_cin" gpurun_out/shape_profile2.log | head
