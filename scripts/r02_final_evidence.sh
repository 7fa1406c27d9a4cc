#!/bin/bash
# round 2, final revision: the evidence copied into profiles/ (1 GPU).  2-GPU lines: scripts/r02_final_2gpu.sh
mkdir -p gpurun_out/f
O=gpurun_out/f
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit,memory.total --format=csv > $O/r02_smi.txt 2>&1; nproc >> $O/r02_smi.txt
exp() { ncu -i $O/$1.ncu-rep --page raw --csv > $O/$1.csv 2>/dev/null; ncu -i $O/$1.ncu-rep --page source --csv > $O/$1.source.csv 2>/dev/null; rm -f $O/$1.ncu-rep; }
timeout 1200 python -m pytest tests -m gpu -q > $O/r02_tests.log 2>&1; echo "tests rc=$?"; tail -n 2 $O/r02_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
timeout 1200 python bench.py --steps 10 --warmup 3 > $O/r02_bench_default.json 2> $O/r02_bench_default.err; echo "bench rc=$?"
python -c "import json; d=json.load(open('$O/r02_bench_default.json')); print({k: d[k] for k in ('value','ms_per_step','unet_fwd_ms','unet_fwd_ms_per_video_at_batch','gpu_launches','clocks')}); print('e2e', d['e2e']); print('roofline', d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['traffic']); print('cpu', d['cpu_baseline']['value'])"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $O/r02_bench_reference_arm.json 2> $O/r02_bench_reference_arm.err; echo "ref arm rc=$?"; cut -c1-300 $O/r02_bench_reference_arm.json
for bs in 1 2 4 8; do
  timeout 500 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --batch $bs > $O/r02_bench_bs$bs.json 2> /dev/null
  python -c "import json; d=json.load(open('$O/r02_bench_bs$bs.json')); print('bs$bs', d['value'], d['e2e']['value'], d['unet_fwd_ms'], d['unet_fwd_ms_per_video_at_batch'], d['roofline']['frac'], d['clocks'])"
done
timeout 600 python bench.py --workload train-step --steps 5 --warmup 2 > $O/r02_train_step_1gpu.json 2> $O/r02_train_step.err; echo "train rc=$?"; python -c "import json; d=json.load(open('$O/r02_train_step_1gpu.json')); print(d['value'], d['ms_per_step'], d['phases'])"
timeout 300 python scripts/train_profile.py families > $O/r02_train_families.txt 2> /dev/null; head -n 12 $O/r02_train_families.txt
timeout 300 python scripts/attn_ablate.py "" p0 p2 p4 p7 q3 q5 a1 a3 a4 a5 v1 > $O/r02_attn_ablate.txt 2>&1; cat $O/r02_attn_ablate.txt
timeout 120 python scripts/attn_bench.py > $O/r02_attn_bench.txt 2>&1; cat $O/r02_attn_bench.txt
timeout 300 python scripts/gemm_bench.py > $O/r02_gemm_shapes.txt 2>&1; tail -n 2 $O/r02_gemm_shapes.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_fwd2 -s 1 -c 1 -o $O/r02_prof_attn2_l0 -f python scripts/attn_bench.py self_l0 > $O/r02_ncu_attn2.log 2>&1; echo "ncu attn2 rc=$?"; exp r02_prof_attn2_l0
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_kernel -s 1 -c 1 -o $O/r02_prof_attn1_l1 -f python scripts/attn_bench.py self_l1 > $O/r02_ncu_attn1.log 2>&1; echo "ncu attn1 rc=$?"; exp r02_prof_attn1_l1
for at in naive sdpa; do
  timeout 400 python scripts/ref_gpu_bench.py --attn $at --batch 8 --calls 3 > $O/r02_ref_gpu_bs8_$at.json 2> $O/r02_ref_gpu_bs8_$at.err; echo "ref gpu $at rc=$?"; cut -c1-400 $O/r02_ref_gpu_bs8_$at.json
done
timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file $O/r02_launches_step.csv python scripts/profile_step.py pipeline > $O/r02_ncu_step.log 2>&1; echo "launch list rc=$?"; wc -l $O/r02_launches_step.csv
du -sh $O
