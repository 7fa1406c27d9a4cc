#!/bin/bash
# final verification + evidence of the round: full GPU test suite, smoke, default bench, reference arm, ncu launch list
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -n 4
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_final.json')); print({k: d[k] for k in ('value','ms_per_step','unet_fwd_ms','gpu_launches','clocks')}); print('e2e', d['e2e']); print('roofline', d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['traffic']); print('cpu', d['cpu_baseline']['value'])"; tail -n 2 gpurun_out/bench_final.err
timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_step_final.csv python scripts/profile_step.py pipeline > gpurun_out/ncu_step_final.log 2>&1; echo "launch list rc=$?"; wc -l gpurun_out/launches_step_final.csv
