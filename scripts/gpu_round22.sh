#!/bin/bash
mkdir -p gpurun_out
( time timeout 1200 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err ) 2>&1 | grep real; echo "default rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_default.json')); print({k: d[k] for k in ('value','ms_per_step','e2e','gpu_launches','cpu_baseline','clocks')}); print(d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['traffic'])"; tail -n 3 gpurun_out/bench_default.err
( time timeout 1200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref2.json 2> gpurun_out/bench_ref2.err ) 2>&1 | grep real; cat gpurun_out/bench_ref2.json; tail -n 3 gpurun_out/bench_ref2.err
