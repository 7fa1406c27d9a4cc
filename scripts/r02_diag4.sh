#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "attention" > gpurun_out/t4_attn.log 2>&1; echo "attention tests rc=$?"; tail -n 4 gpurun_out/t4_attn.log
echo "== attention, two-tile kernel, 16 softmax warps"; timeout 200 python scripts/attn_bench.py 2>&1 | tail -n 6
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 3 -c 1 -o gpurun_out/prof_attn2b_l0 -f python scripts/attn_bench.py self_l0 > gpurun_out/ncu_attn2b.log 2>&1; echo "ncu attn rc=$?"
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench4.json 2> gpurun_out/bench4.err
python -c "import json; d=json.load(open('gpurun_out/bench4.json')); print('bench', d['value'], d['unet_fwd_ms'], d['clocks'])"
