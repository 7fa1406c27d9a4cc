#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_lora_train_gpu.py tests/test_student_gpu.py -m gpu -q > gpurun_out/b9_tests.log 2>&1; echo "tests rc=$?"; tail -n 2 gpurun_out/b9_tests.log
timeout 120 python scripts/wgrad_bench.py 2>&1 | tee gpurun_out/b9_wgrad.txt | tail -n 8
timeout 600 python bench.py --workload train-step --steps 5 --warmup 2 > gpurun_out/b9_train_step.json 2> gpurun_out/b9_train_step.err; echo "train rc=$?"; python -c "import json; d=json.load(open('gpurun_out/b9_train_step.json')); print(d['value'], d['ms_per_step'], d['phases'], d['loss'])"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_lora_train_gpu.py -m gpu -q -k "dropout or wgrad_kernel" > gpurun_out/b9_memcheck_train.log 2>&1; echo "memcheck train rc=$?"; tail -n 4 gpurun_out/b9_memcheck_train.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_kernels_gpu.py tests/test_bwd_kernels_gpu.py -m gpu -q -k "attention and not 2560" > gpurun_out/b9_memcheck_attn.log 2>&1; echo "memcheck attn rc=$?"; tail -n 4 gpurun_out/b9_memcheck_attn.log
