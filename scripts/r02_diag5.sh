#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "attention" > gpurun_out/t5_attn.log 2>&1; echo "attention tests rc=$?"; tail -n 4 gpurun_out/t5_attn.log
echo "== attention, early S release"; timeout 200 python scripts/attn_bench.py 2>&1 | tail -n 6
timeout 600 python -m pytest tests/test_lora_train_gpu.py -m gpu -q -s > gpurun_out/t5_lora.log 2>&1; echo "lora tests rc=$?"; grep -E "^\[|passed|failed|Error|error" gpurun_out/t5_lora.log | head -40
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 3 -c 1 -o gpurun_out/prof_attn2c_l0 -f python scripts/attn_bench.py self_l0 > gpurun_out/ncu_attn2c.log 2>&1; echo "ncu attn rc=$?"
