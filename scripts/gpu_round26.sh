#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "attention" 2>&1 | tail -n 3
timeout 300 python scripts/attn_bench.py 2>&1 | tail -n 5
timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -q -x 2>&1 | tail -n 2
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench20.json 2> gpurun_out/bench20.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench20.json')); print(d['value'], d['unet_fwd_ms'], {k:v for k,v in d['roofline']['families'].items() if k in ('gemm','attn_fwd','groupnorm')})"; tail -n 2 gpurun_out/bench20.err
