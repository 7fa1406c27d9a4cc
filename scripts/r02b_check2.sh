#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_student_gpu.py tests/test_text_encoder.py -m gpu -q -s > gpurun_out/b2_tests.log 2>&1; echo "tests rc=$?"; tail -n 3 gpurun_out/b2_tests.log; grep "distill small" gpurun_out/b2_tests.log
timeout 200 python scripts/attn_ablate.py "" p2 p3 p4 p5 q3 q4 q5 a1 a3 a4 a5 2>&1 | tee gpurun_out/b2_attn_ablate.txt | tail -n 14
timeout 300 python scripts/train_profile.py families > gpurun_out/b2_train_families.txt 2> gpurun_out/b2_train_families.err; cat gpurun_out/b2_train_families.txt; tail -3 gpurun_out/b2_train_families.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/b2_launches_train.csv python scripts/train_profile.py ncu > gpurun_out/b2_ncu_list.log 2>&1; tail -2 gpurun_out/b2_ncu_list.log; wc -l gpurun_out/b2_launches_train.csv
du -sh gpurun_out
