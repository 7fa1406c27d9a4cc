#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x > gpurun_out/t_kernels.log 2>&1; echo "kernel tests rc=$?"; tail -n 4 gpurun_out/t_kernels.log | cut -c1-300
timeout 300 python scripts/gemm_bench.py > gpurun_out/gemm_bench2.log 2>&1; cat gpurun_out/gemm_bench2.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_unet2.csv python scripts/profile_step.py unet > gpurun_out/ncu_unet2.log 2>&1; echo "ncu rc=$?"
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench3.json 2> gpurun_out/bench3.err; echo "bench rc=$?"; cut -c1-420 gpurun_out/bench3.json; tail -n 3 gpurun_out/bench3.err
