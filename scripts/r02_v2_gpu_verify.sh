#!/bin/bash
# Full GPU verification + first measurement of the v2 full fine-tune step.  The round's last 1.7 GPU-minutes ran
# only the two pytest selections of step 1 (profiles/r02_v2_gpu_*_tests.log); THIS SCRIPT as a whole has not been executed.
# One call, ~5 GPU-minutes:   gpurun --timeout 900 -- 'bash scripts/r02_v2_gpu_verify.sh'
#   1. every test of tests/test_zz_full_train_gpu.py (v2 step, decoder grad, motion-prior score, mid-size student) WITHOUT the xfail mask (T2V_ZZ_CHILD=1 runs the never-run bodies in-process; --runxfail: a failure is a failure here)
#   2. compute-sanitizer memcheck over the three new kernels + wgrad_wide
#   3. the unmeasured bench line (eager, one sample per step) and its launch list
mkdir -p gpurun_out/v2
T2V_ZZ_CHILD=1 timeout 600 python -m pytest tests/test_zz_full_train_gpu.py -m gpu --runxfail -q -s -k "not own_process" > gpurun_out/v2/tests.log 2>&1; echo "v2 tests rc=$?"; tail -n 5 gpurun_out/v2/tests.log
T2V_ZZ_CHILD=1 timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_zz_full_train_gpu.py -m gpu --runxfail -q \
  -k "affine_grad or ema_update or wgrad_wide or softmax_bwd or probs_bwd" > gpurun_out/v2/memcheck.log 2>&1; echo "memcheck rc=$?"; tail -n 3 gpurun_out/v2/memcheck.log
timeout 900 python bench.py --workload v2-step --steps 3 --warmup 2 > gpurun_out/v2/bench_v2_step.json 2> gpurun_out/v2/bench_v2_step.err; echo "bench rc=$?"
tail -c 600 gpurun_out/v2/bench_v2_step.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/v2/launches_v2_step.csv \
  python bench.py --workload v2-step --steps 1 --warmup 2 > /dev/null 2>&1; echo "ncu rc=$?"
python scripts/summarize_launches.py gpurun_out/v2/launches_v2_step.csv > gpurun_out/v2/launches_v2_step_summary.txt 2>&1 || true
