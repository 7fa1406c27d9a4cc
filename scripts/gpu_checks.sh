#!/bin/bash
# Run the GPU kernel parity tests group by group (a trap in one group must not hide the others).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/smi.txt 2>&1
nproc >> gpurun_out/smi.txt; lscpu | grep "Model name" >> gpurun_out/smi.txt
for k in "test_linear or test_geglu" "test_conv or test_tconv or test_bmm" "norm" "test_attention and not temporal" "temporal" "small or layout or lcm"; do
  name=$(echo "$k" | tr ' ' '_')
  timeout 420 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "$k" > "gpurun_out/t_${name}.log" 2>&1
  echo "[$k] rc=$?"
  tail -n 3 "gpurun_out/t_${name}.log"
done
