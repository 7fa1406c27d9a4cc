#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "geglu or pairs" 2>&1 | tail -n 3
timeout 300 python scripts/gemm_bench.py geglu320 geglu640 geglu1280 lin320_res 2>&1 | tail -n 4
timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -q -x -s 2>&1 | grep -E "unet full|passed|failed" | cut -c1-250
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench8.json 2> gpurun_out/bench8.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench8.json')); print(d['value'], d['unet_fwd_ms'], {k:v for k,v in d['roofline']['families'].items() if k in ('gemm','attn_fwd','groupnorm')})"; tail -n 2 gpurun_out/bench8.err
