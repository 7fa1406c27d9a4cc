#!/bin/bash
# whole-warp MMA issuers (gemm_tc, attention fwd / fwd2 / bwd), two-tile attention on long sequences: full suite + lines
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/b5_tests.log 2>&1; echo "tests rc=$?"; tail -n 3 gpurun_out/b5_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
timeout 120 python scripts/attn_bench.py 2>&1 | tee gpurun_out/b5_attn_bench.txt | tail -n 6
timeout 300 python scripts/gemm_bench.py > gpurun_out/b5_gemm_shapes.txt 2>&1; tail -n 30 gpurun_out/b5_gemm_shapes.txt
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/b5_bench_default.json 2> gpurun_out/b5_bench_default.err; echo "bench rc=$?"
python -c "import json; d=json.load(open('gpurun_out/b5_bench_default.json')); print({k: d[k] for k in ('value','ms_per_step','unet_fwd_ms','unet_fwd_ms_per_video_at_batch','gpu_launches','clocks')}); print('e2e', d['e2e']); print('roofline', d['roofline']['achieved'], d['roofline']['frac']); print({k:(v['ms']) for k,v in d['roofline']['families'].items()})"
timeout 600 python bench.py --workload train-step --steps 5 --warmup 2 > gpurun_out/b5_train_step.json 2> gpurun_out/b5_train_step.err; echo "train rc=$?"; python -c "import json; d=json.load(open('gpurun_out/b5_train_step.json')); print(d['value'], d['ms_per_step'], d['phases'])"
