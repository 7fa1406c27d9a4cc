#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_lora_train_gpu.py tests/test_student_gpu.py -m gpu -q > gpurun_out/b6_tests.log 2>&1; echo "tests rc=$?"; tail -n 3 gpurun_out/b6_tests.log
timeout 600 python bench.py --workload train-step --steps 5 --warmup 2 > gpurun_out/b6_train_step.json 2> gpurun_out/b6_train_step.err; echo "train rc=$?"; python -c "import json; d=json.load(open('gpurun_out/b6_train_step.json')); print(d['value'], d['ms_per_step'], d['phases'], d['loss'])"; tail -n 3 gpurun_out/b6_train_step.err
for bs in 12 16; do
  timeout 500 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --batch $bs > gpurun_out/b6_bench_bs$bs.json 2> gpurun_out/b6_bench_bs$bs.err
  python -c "import json; d=json.load(open('gpurun_out/b6_bench_bs$bs.json')); print('bs$bs', d['value'], d['e2e']['value'], d['unet_fwd_ms_per_video_at_batch'], d['roofline']['frac'], d['clocks'])" || tail -n 3 gpurun_out/b6_bench_bs$bs.err
done
