#!/bin/bash
# final verification of the round: full GPU test suite, smoke, default bench (add the ncu launch list with scripts/gpu_evidence.sh)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -n 4
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_final.json')); print({k: d[k] for k in ('value','ms_per_step','unet_fwd_ms','gpu_launches','clocks')}); print('e2e', d['e2e']); print('roofline', d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['traffic']); print('cpu', d['cpu_baseline']['value'])"; tail -n 2 gpurun_out/bench_final.err
