#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_lora_train_gpu.py tests/test_student_gpu.py -m gpu -q -s > gpurun_out/b7_tests.log 2>&1; echo "tests rc=$?"; tail -n 3 gpurun_out/b7_tests.log; grep "distill small\|student" gpurun_out/b7_tests.log | head
timeout 600 python bench.py --workload train-step --steps 5 --warmup 2 > gpurun_out/b7_train_step.json 2> gpurun_out/b7_train_step.err; echo "train rc=$?"; python -c "import json; d=json.load(open('gpurun_out/b7_train_step.json')); print(d['value'], d['ms_per_step'], d['phases'], d['loss'])"; tail -n 3 gpurun_out/b7_train_step.err
