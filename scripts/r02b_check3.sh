#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_bwd_kernels_gpu.py -m gpu -q -k "attention" > gpurun_out/b3_tests_v1.log 2>&1; echo "tests v1 rc=$?"; tail -n 2 gpurun_out/b3_tests_v1.log
for var in p2 q3; do
T2V_ATTN_V2=1 T2V_ATTN_V2_VARIANT=$var timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "attention and not temporal" > gpurun_out/b3_tests_v2_$var.log 2>&1; echo "tests v2 $var rc=$?"; tail -n 2 gpurun_out/b3_tests_v2_$var.log
done
timeout 300 python scripts/attn_ablate.py "" p0 p2 p3 p4 p7 q3 q5 q7 a1 a3 v1 2>&1 | tee gpurun_out/b3_attn_ablate.txt | tail -n 14
timeout 120 python scripts/attn_bench.py 2>&1 | tee gpurun_out/b3_attn_bench_v1.txt | tail -n 6
T2V_ATTN_V2=1 T2V_ATTN_V2_VARIANT=p2 timeout 120 python scripts/attn_bench.py 2>&1 | tee gpurun_out/b3_attn_bench_v2.txt | tail -n 6
