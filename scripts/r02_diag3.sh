#!/bin/bash
# round 2, GPU call 3: two-tile attention kernel (parity, A/B timing, ncu), GroupNorm statistics (per-warp tables), no zero kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "attention" > gpurun_out/t3_attn.log 2>&1; echo "attention tests rc=$?"; tail -n 4 gpurun_out/t3_attn.log
echo "== attention, two-tile kernel"; timeout 200 python scripts/attn_bench.py 2>&1 | tail -n 6
echo "== attention, single-tile kernel (round 1)"; T2V_ATTN_V1=1 timeout 200 python scripts/attn_bench.py 2>&1 | tail -n 6
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/tests3.log 2>&1; echo "all tests rc=$?"; tail -n 4 gpurun_out/tests3.log
for st in 0 1; do STATS=$st timeout 200 python scripts/gemm_bench.py conv320 tconv320 vae256 vae128 conv1280_l3 > gpurun_out/gemm3_stats$st.txt 2>&1; echo "== STATS=$st"; cat gpurun_out/gemm3_stats$st.txt; done
for mode in "conv 2304" "conv 1152" "all 0" "off 0"; do set -- $mode
  T2V_GN_FUSE=$1 T2V_GN_FUSE_MIN_K=$2 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench3_$1_$2.json 2> gpurun_out/bench3_$1_$2.err
  python -c "import json; d=json.load(open('gpurun_out/bench3_$1_$2.json')); print('gn $1 $2', d['value'], d['unet_fwd_ms'], d['roofline']['frac'], d['clocks']['sm_mhz'])"
done
T2V_GEMM_TUNE=0x1000000 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench3_zero.json 2> gpurun_out/bench3_zero.err
python -c "import json; d=json.load(open('gpurun_out/bench3_zero.json')); print('with zero kernel', d['value'], d['unet_fwd_ms'])"
T2V_ATTN_V1=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench3_attnv1.json 2> gpurun_out/bench3_attnv1.err
python -c "import json; d=json.load(open('gpurun_out/bench3_attnv1.json')); print('attention v1', d['value'], d['unet_fwd_ms'])"
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --batch 8 > gpurun_out/bench3_bs8.json 2> gpurun_out/bench3_bs8.err
python -c "import json; d=json.load(open('gpurun_out/bench3_bs8.json')); print('bs8', d['value'], d['e2e']['value'], d['unet_fwd_ms'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 3 -c 1 -o gpurun_out/prof_attn2_l0 -f python scripts/attn_bench.py self_l0 > gpurun_out/ncu_attn2.log 2>&1; echo "ncu attn rc=$?"
