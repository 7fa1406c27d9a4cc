#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ms.py -m gpu -q -s > gpurun_out/t7_ms.log 2>&1; echo "ms tests rc=$?"; grep -E "^\[ms|passed|failed|Error" gpurun_out/t7_ms.log | head -20
timeout 600 python bench.py --workload lora-step --steps 5 --warmup 2 > gpurun_out/lora_step_1gpu.json 2> gpurun_out/lora_step_1gpu.err; echo "lora-step rc=$?"; cat gpurun_out/lora_step_1gpu.json; tail -n 4 gpurun_out/lora_step_1gpu.err
