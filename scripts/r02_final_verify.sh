#!/bin/bash
# last revision: full GPU suite, smoke, the default bench command
mkdir -p gpurun_out/v
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/v/tests.log 2>&1; echo "tests rc=$?"; tail -n 2 gpurun_out/v/tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
timeout 600 python bench.py --steps 4 --warmup 3 > gpurun_out/v/bench_default.json 2> gpurun_out/v/bench_default.err; echo "bench rc=$?"
python -c "import json; d=json.load(open('gpurun_out/v/bench_default.json')); r=d['roofline']; print({k: d[k] for k in ('metric','value','ms_per_step','unet_fwd_ms','gpu_launches','clocks')}); print('e2e', d['e2e']); print('roofline', r['achieved'], r['frac'], r['traffic'], r['traffic_batch'], r['algorithmic_bytes_per_launch']); print('cpu', d['cpu_baseline']['value'])" || tail -n 5 gpurun_out/v/bench_default.err
