#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x > gpurun_out/t_kernels.log 2>&1; echo "kernel tests rc=$?"; tail -n 3 gpurun_out/t_kernels.log | cut -c1-300
T2V_GEMM_TUNE=0 timeout 300 python scripts/gemm_bench.py lin320_res lin320 qkv320 geglu320 ff2_320 lin640_res geglu640 geglu1280 lin1280_res conv320 conv640 tconv320 vae512 2>&1 | tail -n 13
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench6.json 2> gpurun_out/bench6.err; echo "bench rc=$?"; cut -c1-330 gpurun_out/bench6.json; tail -n 2 gpurun_out/bench6.err
