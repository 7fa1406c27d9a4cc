#!/bin/bash
# round 2, GPU call 2: tests with observed tolerances, split-K fix-up A/B, GroupNorm-statistics epilogue cost per shape, batch sweep
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/tests2.log 2>&1; echo "tests rc=$?"; tail -n 6 gpurun_out/tests2.log
for st in 0 1; do STATS=$st timeout 200 python scripts/gemm_bench.py conv320 conv640 tconv320 vae512 vae256 vae128 > gpurun_out/gemm_stats$st.txt 2>&1; echo "== STATS=$st"; cat gpurun_out/gemm_stats$st.txt; done
timeout 200 python scripts/gemm_bench.py lin1280_l3 conv1280_l3 tconv1280_l3 conv1280 > gpurun_out/gemm_l3_fix.txt 2>&1; echo "== fixup"; cat gpurun_out/gemm_l3_fix.txt
T2V_GEMM_TUNE=0x1000000 timeout 200 python scripts/gemm_bench.py lin1280_l3 conv1280_l3 tconv1280_l3 conv1280 > gpurun_out/gemm_l3_legacy.txt 2>&1; echo "== legacy"; cat gpurun_out/gemm_l3_legacy.txt
timeout 200 python scripts/gn_bench.py > gpurun_out/gn_bench.txt 2>&1; cat gpurun_out/gn_bench.txt
for bs in 1 2 4; do
  timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --batch $bs > gpurun_out/bench2_bs$bs.json 2> gpurun_out/bench2_bs$bs.err; echo "bench bs=$bs rc=$?"
  python -c "import json; d=json.load(open('gpurun_out/bench2_bs$bs.json')); print('bs$bs', d['value'], d['e2e']['value'], d['unet_fwd_ms'], d['roofline']['frac'], d['gpu_launches'], d['clocks'])"
done
T2V_GEMM_TUNE=0x1000000 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench2_legacy.json 2> gpurun_out/bench2_legacy.err
python -c "import json; d=json.load(open('gpurun_out/bench2_legacy.json')); print('legacy split-K', d['value'], d['unet_fwd_ms'], d['gpu_launches'])"
