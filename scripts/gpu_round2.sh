#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -q -s > gpurun_out/t_model.log 2>&1; echo "model tests rc=$?"; grep -E "^\[|passed|failed|Error|error" gpurun_out/t_model.log | head -40
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench1.json 2> gpurun_out/bench1.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench1.json; tail -n 5 gpurun_out/bench1.err
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cat gpurun_out/bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_unet.csv python scripts/profile_step.py unet > gpurun_out/ncu_unet.log 2>&1; echo "ncu rc=$?"; tail -n 2 gpurun_out/ncu_unet.log; wc -l gpurun_out/launches_unet.csv
