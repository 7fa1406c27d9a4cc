#!/bin/bash
# round 2, GPU call 1: full GPU test suite (observed errors printed), per-shape profiles at bs 1 / 2, bench with the
# GroupNorm-statistics fusion on / off and at bs 2, the unmodified reference on the GPU (naive / SDPA) and on the host cores
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/smi.txt 2>&1
nproc >> gpurun_out/smi.txt; lscpu | grep -E "Model name|Flags" | head -2 | cut -c1-400 >> gpurun_out/smi.txt
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/tests.log 2>&1; echo "tests rc=$?"; tail -n 5 gpurun_out/tests.log
timeout 300 python scripts/shape_profile.py 1 200 > gpurun_out/shape_bs1.txt 2>&1; echo "shape1 rc=$?"
timeout 300 python scripts/shape_profile.py 2 200 > gpurun_out/shape_bs2.txt 2>&1; echo "shape2 rc=$?"
for mode in conv off all; do
  T2V_GN_FUSE=$mode timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_gn_$mode.json 2> gpurun_out/bench_gn_$mode.err; echo "bench gn=$mode rc=$?"
  python -c "import json; d=json.load(open('gpurun_out/bench_gn_$mode.json')); print('$mode', d['value'], d['unet_fwd_ms'], d['roofline']['frac'], d['clocks'])"
done
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --batch 2 > gpurun_out/bench_bs2.json 2> gpurun_out/bench_bs2.err; echo "bench bs2 rc=$?"
python -c "import json; d=json.load(open('gpurun_out/bench_bs2.json')); print('bs2', d['value'], d['unet_fwd_ms'], d['roofline']['frac'])"
timeout 400 python scripts/ref_gpu_bench.py --attn naive > gpurun_out/ref_gpu_naive.json 2> gpurun_out/ref_gpu_naive.err; echo "ref naive rc=$?"; cat gpurun_out/ref_gpu_naive.json
timeout 400 python scripts/ref_gpu_bench.py --attn sdpa > gpurun_out/ref_gpu_sdpa.json 2> gpurun_out/ref_gpu_sdpa.err; echo "ref sdpa rc=$?"; cat gpurun_out/ref_gpu_sdpa.json
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/ref_cpu.json 2> gpurun_out/ref_cpu.err; echo "ref cpu rc=$?"; cat gpurun_out/ref_cpu.json
