#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x 2>&1 | tail -n 6
timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -q -x -s 2>&1 | grep -E "passed|failed|rel_l2|Error|error" | tail -n 12
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench16.json 2> gpurun_out/bench16.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench16.json')); print(d['value'], d['unet_fwd_ms'], {k:v for k,v in d['roofline']['families'].items() if k in ('gemm','attn_fwd','groupnorm','layernorm')})"; tail -n 2 gpurun_out/bench16.err
timeout 300 python scripts/gemm_bench.py qkv320 geglu320 geglu640 2>&1 | tail -n 3
