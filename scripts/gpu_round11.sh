#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x > gpurun_out/t_kernels.log 2>&1; echo "kernel tests rc=$?"; tail -n 3 gpurun_out/t_kernels.log | cut -c1-300
timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -q -x > gpurun_out/t_model.log 2>&1; echo "model tests rc=$?"; tail -n 3 gpurun_out/t_model.log | cut -c1-300
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench5.json 2> gpurun_out/bench5.err; echo "bench rc=$?"; cut -c1-330 gpurun_out/bench5.json; tail -n 2 gpurun_out/bench5.err
timeout 300 python scripts/ref_gpu_bench.py > gpurun_out/ref_gpu.json 2> gpurun_out/ref_gpu.err; echo "ref gpu rc=$?"; cat gpurun_out/ref_gpu.json; tail -n 2 gpurun_out/ref_gpu.err
