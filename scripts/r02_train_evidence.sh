#!/bin/bash
# training-step evidence: tests, the 1-GPU train-step line, the unmodified reference student on the same GPU
mkdir -p gpurun_out/d17
timeout 300 python -m pytest tests/test_student_gpu.py -q -m gpu -s 2>&1 | tail -25 > gpurun_out/d17/tests.txt; cat gpurun_out/d17/tests.txt
timeout 400 python bench.py --workload train-step --steps 5 --warmup 2 > gpurun_out/d17/train_step.json 2> gpurun_out/d17/train_step.err; echo "train-step rc=$?"; cat gpurun_out/d17/train_step.json
timeout 400 python scripts/ref_gpu_train.py --attn sdpa > gpurun_out/d17/ref_train_sdpa.json 2> gpurun_out/d17/ref_train_sdpa.err; echo "ref sdpa rc=$?"; tail -1 gpurun_out/d17/ref_train_sdpa.json; tail -3 gpurun_out/d17/ref_train_sdpa.err
timeout 400 python scripts/ref_gpu_train.py --attn naive > gpurun_out/d17/ref_train_naive.json 2> gpurun_out/d17/ref_train_naive.err; echo "ref naive rc=$?"; tail -1 gpurun_out/d17/ref_train_naive.json; tail -3 gpurun_out/d17/ref_train_naive.err
