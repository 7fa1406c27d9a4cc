#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x 2>&1 | tail -n 3
timeout 300 python scripts/gemm_bench.py lin1280_l3 conv1280_l3 tconv1280_l3 2>&1 | tail -n 3
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench17.json 2> gpurun_out/bench17.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench17.json')); print(d['value'], d['unet_fwd_ms'], {k:v for k,v in d['roofline']['families'].items() if k in ('gemm','attn_fwd','groupnorm','layernorm')})"; tail -n 2 gpurun_out/bench17.err
