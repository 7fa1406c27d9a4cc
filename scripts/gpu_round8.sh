#!/bin/bash
mkdir -p gpurun_out
CASES="lin320_res lin320 geglu320 lin640_res lin1280_res conv320"
for tune in 0 0x1000 0x2000 0x3000 0x7000; do
  echo "== T2V_GEMM_TUNE=$tune"
  T2V_GEMM_TUNE=$tune timeout 300 python scripts/gemm_bench.py $CASES 2>&1 | tail -n 6
done
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench4.json 2> gpurun_out/bench4.err; echo "bench rc=$?"; cut -c1-330 gpurun_out/bench4.json; tail -n 2 gpurun_out/bench4.err
