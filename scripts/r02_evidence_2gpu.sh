#!/bin/bash
# 2-GPU lines: inference replicas and the LoRA-step exchange (NCCL all-reduce of the gradient arena)
mkdir -p gpurun_out
timeout 600 python bench.py --workload lora-step --steps 5 --warmup 2 > gpurun_out/r02_lora_step_1gpu_inbox.json 2> gpurun_out/r02_lora_step_1gpu_inbox.err; echo "lora-step 1gpu rc=$?"; cat gpurun_out/r02_lora_step_1gpu_inbox.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --workload lora-step --gpus 2 --steps 5 --warmup 2 > gpurun_out/r02_lora_step_2gpu.json 2> gpurun_out/r02_lora_step_2gpu.err; echo "lora-step 2gpu rc=$?"; cat gpurun_out/r02_lora_step_2gpu.json; tail -n 3 gpurun_out/r02_lora_step_2gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err; echo "bench 2gpu rc=$?"; python -c "import json; d=json.load(open('gpurun_out/r02_bench_2gpu.json')); print(d['value'], d['n_gpus'], d['e2e'])"
