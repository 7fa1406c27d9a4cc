#!/bin/bash
# round 2 evidence for profiles/: full GPU suite + smoke, the default bench (both arms), batch sweep, ncu launch list of one
# default step, --set full captures of the dominant kernels.  (The 2-GPU lines come from scripts/r02_evidence_2gpu.sh.)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit,memory.total --format=csv > gpurun_out/r02_smi.txt 2>&1; nproc >> gpurun_out/r02_smi.txt
# (.ncu-rep files are exported to --page raw csv on the box and deleted: gpurun merges at most 64 MiB back)
exp() { ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1.csv 2>/dev/null; ncu -i gpurun_out/$1.ncu-rep --page source --csv > gpurun_out/$1.source.csv 2>/dev/null; rm -f gpurun_out/$1.ncu-rep; }
timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/r02_tests.log 2>&1; echo "tests rc=$?"; tail -n 3 gpurun_out/r02_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; echo "bench rc=$?"
python -c "import json; d=json.load(open('gpurun_out/r02_bench_default.json')); print({k: d[k] for k in ('value','ms_per_step','unet_fwd_ms','unet_fwd_ms_per_video_at_batch','gpu_launches','clocks')}); print('e2e', d['e2e']); print('roofline', d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['traffic']); print('cpu', d['cpu_baseline'])"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err; echo "ref arm rc=$?"; cut -c1-400 gpurun_out/r02_bench_reference_arm.json
for bs in 1 2 4; do
  timeout 500 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --batch $bs > gpurun_out/r02_bench_bs$bs.json 2> /dev/null
  python -c "import json; d=json.load(open('gpurun_out/r02_bench_bs$bs.json')); print('bs$bs', d['value'], d['e2e']['value'], d['unet_fwd_ms'], d['roofline']['frac'], d['clocks'])"
done
timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_step.csv python scripts/profile_step.py pipeline > gpurun_out/r02_ncu_step.log 2>&1; echo "launch list rc=$?"; wc -l gpurun_out/r02_launches_step.csv
for c in vae512 vae128 conv320 lin320_res geglu320; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 1 -c 1 -o gpurun_out/r02_prof_gemm_$c -f python scripts/gemm_bench.py $c > gpurun_out/r02_ncu_gemm_$c.log 2>&1; echo "ncu gemm $c rc=$?"; exp r02_prof_gemm_$c
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_kernel -s 1 -c 1 -o gpurun_out/r02_prof_attn_v1_l0 -f python scripts/attn_bench.py self_l0 > gpurun_out/r02_ncu_attn1.log 2>&1; echo "ncu attn v1 rc=$?"; exp r02_prof_attn_v1_l0
timeout 300 ncu --set full --clock-control none -k regex:gn_stats -s 1 -c 1 -o gpurun_out/r02_prof_gn_stats -f python scripts/gn_bench.py big > gpurun_out/r02_ncu_gn1.log 2>&1; echo "ncu gn_stats rc=$?"; exp r02_prof_gn_stats
timeout 300 ncu --set full --clock-control none -k regex:gn_apply -s 1 -c 1 -o gpurun_out/r02_prof_gn_apply -f python scripts/gn_bench.py big > gpurun_out/r02_ncu_gn2.log 2>&1; echo "ncu gn_apply rc=$?"; exp r02_prof_gn_apply
timeout 300 ncu --set full --clock-control none -k regex:wgrad_tc -s 2 -c 1 -o gpurun_out/r02_prof_wgrad -f python scripts/wgrad_bench.py > gpurun_out/r02_ncu_wgrad.log 2>&1; echo "ncu wgrad rc=$?"; exp r02_prof_wgrad
timeout 120 python scripts/wgrad_bench.py 2>&1 | tail -n 5
T2V_ATTN_V2=1 timeout 120 python scripts/attn_bench.py 2>&1 | tail -n 5
timeout 300 python scripts/gemm_bench.py > gpurun_out/r02_gemm_shapes.txt 2>&1; tail -n 3 gpurun_out/r02_gemm_shapes.txt
timeout 200 python scripts/gn_bench.py > gpurun_out/r02_gn_bench.txt 2>&1; timeout 100 python scripts/gn_bench.py big >> gpurun_out/r02_gn_bench.txt 2>&1; tail -n 3 gpurun_out/r02_gn_bench.txt
du -sh gpurun_out
