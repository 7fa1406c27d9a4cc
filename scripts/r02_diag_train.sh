#!/bin/bash
mkdir -p gpurun_out/d19
timeout 300 python scripts/train_profile.py families > gpurun_out/d19/families.txt 2> gpurun_out/d19/families.err; cat gpurun_out/d19/families.txt; tail -3 gpurun_out/d19/families.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/d19/launches_train.csv python scripts/train_profile.py ncu > gpurun_out/d19/ncu_list.log 2>&1; tail -2 gpurun_out/d19/ncu_list.log; wc -l gpurun_out/d19/launches_train.csv
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:attn_bwd_kernel -c 2 -o gpurun_out/d19/attn_bwd python scripts/train_profile.py ncu > gpurun_out/d19/ncu_attn_bwd.log 2>&1; tail -2 gpurun_out/d19/ncu_attn_bwd.log
ncu -i gpurun_out/d19/attn_bwd.ncu-rep --page raw --csv > gpurun_out/d19/r02_ncu_full_attn_bwd.csv 2>/dev/null; rm -f gpurun_out/d19/attn_bwd.ncu-rep; wc -c gpurun_out/d19/r02_ncu_full_attn_bwd.csv
