#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "small_cin" 2>&1 | tail -n 2
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/d8_bench.json 2> gpurun_out/d8_bench.err; python -c "import json; d=json.load(open('gpurun_out/d8_bench.json')); print('bench', d['value'], d['e2e']['value'], d['unet_fwd_ms'], d['roofline']['frac'], d['clocks']); print({k: v for k, v in list(d['roofline']['families'].items())[:6]})"
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --batch 1 > gpurun_out/d8_bench_bs1.json 2> /dev/null; python -c "import json; d=json.load(open('gpurun_out/d8_bench_bs1.json')); print('bs1', d['value'], d['unet_fwd_ms'], d['clocks'])"
timeout 2400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_step.csv python scripts/profile_step.py pipeline > gpurun_out/r02_ncu_step.log 2>&1; echo "launch list rc=$?"; wc -l gpurun_out/r02_launches_step.csv; tail -n 1 gpurun_out/r02_ncu_step.log
