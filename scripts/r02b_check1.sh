#!/bin/bash
# session-2 state check: full GPU suite, smoke, default bench, train-step line, attention variants + one source-level capture
mkdir -p gpurun_out
exp() { ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1.csv 2>/dev/null; ncu -i gpurun_out/$1.ncu-rep --page source --csv > gpurun_out/$1.source.csv 2>/dev/null; rm -f gpurun_out/$1.ncu-rep; }
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/b1_tests.log 2>&1; echo "tests rc=$?"; tail -n 3 gpurun_out/b1_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
timeout 200 python scripts/attn_ablate.py 2>&1 | tee gpurun_out/b1_attn_ablate.txt | tail -n 12
T2V_ATTN_V2=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_fwd2 -s 1 -c 1 -o gpurun_out/b1_prof_attn2 -f python scripts/attn_bench.py self_l0 > gpurun_out/b1_ncu_attn2.log 2>&1; echo "ncu attn2 rc=$?"; exp b1_prof_attn2
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/b1_bench_default.json 2> gpurun_out/b1_bench_default.err; echo "bench rc=$?"
python -c "import json; d=json.load(open('gpurun_out/b1_bench_default.json')); print({k: d[k] for k in ('value','ms_per_step','unet_fwd_ms','unet_fwd_ms_per_video_at_batch','gpu_launches','clocks')}); print('e2e', d['e2e']); print('roofline', d['roofline']['achieved'], d['roofline']['frac']); print({k:(v['ms']) for k,v in d['roofline']['families'].items()})"
timeout 600 python bench.py --workload train-step --steps 5 --warmup 2 > gpurun_out/b1_train_step.json 2> gpurun_out/b1_train_step.err; echo "train rc=$?"; cut -c1-1500 gpurun_out/b1_train_step.json
du -sh gpurun_out
