#!/bin/bash
mkdir -p gpurun_out
export T2V_PDL=0
echo skip
CASES="lin320_res lin320 qkv320 geglu320 ff2_320 lin640_res geglu640 lin1280_res conv320 tconv320"
for tune in 0 0x200; do
  echo "== T2V_GEMM_TUNE=$tune"
  T2V_GEMM_TUNE=$tune timeout 300 python scripts/gemm_bench.py $CASES 2>&1 | tail -n 10
done
export T2V_PDL=1
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x > gpurun_out/t_kernels_pdl.log 2>&1; echo "kernel tests (PDL on) rc=$?"; tail -n 4 gpurun_out/t_kernels_pdl.log | cut -c1-400
timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -q -x 2>&1 | tail -n 3
for pdl in 0 1; do
  T2V_PDL=$pdl timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pdl$pdl.json 2> gpurun_out/bench_pdl$pdl.err; echo "bench PDL=$pdl rc=$?"; cut -c1-330 gpurun_out/bench_pdl$pdl.json; tail -n 2 gpurun_out/bench_pdl$pdl.err
done
