#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x > gpurun_out/t_kernels.log 2>&1; echo "kernel tests rc=$?"; tail -n 15 gpurun_out/t_kernels.log | cut -c1-400
timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -q -x -s > gpurun_out/t_model.log 2>&1; echo "model tests rc=$?"; grep -E "^\[|passed|failed|Error" gpurun_out/t_model.log | cut -c1-300 | head -20
timeout 300 python scripts/shape_profile.py > gpurun_out/shape_profile.log 2>&1; echo "shape rc=$?"; grep "==" gpurun_out/shape_profile.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench2.json 2> gpurun_out/bench2.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/bench2.json; tail -n 3 gpurun_out/bench2.err
