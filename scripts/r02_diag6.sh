#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/tests6.log 2>&1; echo "all tests rc=$?"; tail -n 5 gpurun_out/tests6.log; grep -E "^\[attention probs|^\[lora" gpurun_out/tests6.log | head
timeout 600 python bench.py --workload lora-step --steps 5 --warmup 2 > gpurun_out/lora_step_1gpu.json 2> gpurun_out/lora_step_1gpu.err; echo "lora-step rc=$?"; cat gpurun_out/lora_step_1gpu.json; tail -n 5 gpurun_out/lora_step_1gpu.err
for bs in 1 4 8; do
  timeout 500 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --batch $bs > gpurun_out/bench6_bs$bs.json 2> gpurun_out/bench6_bs$bs.err
  python -c "import json; d=json.load(open('gpurun_out/bench6_bs$bs.json')); print('bs$bs', d['value'], d['e2e']['value'], d['unet_fwd_ms'], d['roofline']['frac'], d['clocks'])"
done
