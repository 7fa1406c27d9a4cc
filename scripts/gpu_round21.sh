#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "groupnorm" 2>&1 | tail -n 4
echo "== T2V_GN_FUSE=all"
T2V_GN_FUSE=all timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -q -x -s 2>&1 | grep -E "passed|failed|AssertionError: |Error|unet full" | head -5
for pol in off conv; do
echo "== policy $pol"
T2V_GN_FUSE=$pol timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_gn_$pol.json 2> gpurun_out/bench_gn_$pol.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_gn_$pol.json')); print(d['value'], d['unet_fwd_ms'], {k:v for k,v in d['roofline']['families'].items() if k in ('gemm','attn_fwd','groupnorm','layernorm')})"; tail -n 2 gpurun_out/bench_gn_$pol.err
done
T2V_GN_FUSE=conv T2V_GN_FUSE_MIN_K=2304 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('min_k 2304:', d['value'], d['unet_fwd_ms'], {k:v for k,v in d['roofline']['families'].items() if k in ('gemm','groupnorm')})"
