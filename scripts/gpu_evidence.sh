#!/bin/bash
# ncu evidence for profiles/: launch list (+DRAM bytes) of ONE pipeline step, --set full captures of the top kernels
mkdir -p gpurun_out
timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_step.csv python scripts/profile_step.py pipeline > gpurun_out/ncu_step.log 2>&1; echo "launch list rc=$?"; wc -l gpurun_out/launches_step.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 1 -c 1 -o gpurun_out/prof_gemm_vae512 -f python scripts/gemm_bench.py vae512 > gpurun_out/ncu_gemm1.log 2>&1; echo "gemm vae512 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 1 -c 1 -o gpurun_out/prof_gemm_conv320 -f python scripts/gemm_bench.py conv320 > gpurun_out/ncu_gemm2.log 2>&1; echo "gemm conv320 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 1 -c 1 -o gpurun_out/prof_gemm_lin320 -f python scripts/gemm_bench.py lin320_res > gpurun_out/ncu_gemm3.log 2>&1; echo "gemm lin320 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 1 -c 1 -o gpurun_out/prof_attn_l0 -f python scripts/attn_bench.py self_l0 > gpurun_out/ncu_attn.log 2>&1; echo "attn rc=$?"
timeout 300 python scripts/attn_bench.py 2>&1 | tail -n 6
