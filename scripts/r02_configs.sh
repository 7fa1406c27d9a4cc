#!/bin/bash
# BASELINE configs[2] (8- / 16-step v2 path with motion conditioning) and configs[4] (ModelScope 4-step 16x256x256) lines
mkdir -p gpurun_out/c
O=gpurun_out/c
show() { python -c "import json,sys; d=json.load(open('$1')); print(d['metric'], '|', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms/step', round(d['ms_per_step'],1), 'gemm frac', round(d['roofline']['frac'],3), d['clocks']['sm_mhz'], d['config']['workload'][:60])" || tail -n 5 ${1%.json}.err; }
timeout 600 python bench.py --sample-steps 8 --motion-cond --batch 8 --steps 3 --warmup 3 --no-cpu-baseline > $O/r02_bench_8step_motion.json 2> $O/r02_bench_8step_motion.err; show $O/r02_bench_8step_motion.json
timeout 600 python bench.py --sample-steps 16 --motion-cond --batch 8 --steps 3 --warmup 3 --no-cpu-baseline > $O/r02_bench_16step_motion.json 2> $O/r02_bench_16step_motion.err; show $O/r02_bench_16step_motion.json
timeout 600 python bench.py --workload ms-pipeline --batch 16 --steps 5 --warmup 3 --no-cpu-baseline > $O/r02_bench_ms_bs16.json 2> $O/r02_bench_ms_bs16.err; show $O/r02_bench_ms_bs16.json
timeout 600 python bench.py --workload ms-pipeline --batch 1 --steps 10 --warmup 3 --no-cpu-baseline > $O/r02_bench_ms_bs1.json 2> $O/r02_bench_ms_bs1.err; show $O/r02_bench_ms_bs1.json
