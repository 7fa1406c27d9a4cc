#!/bin/bash
# the distillation step on N GPUs of one box (config 4: DDP over 8 x B200 with the NCCL LoRA-gradient all-reduce)
N=${1:-8}
mkdir -p gpurun_out/t8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --workload train-step --gpus $N --steps 5 --warmup 2 > gpurun_out/t8/r02_train_step_${N}gpu.json 2> gpurun_out/t8/r02_train_step_${N}gpu.err; echo "train-step ${N}gpu rc=$?"; python -c "import json; d=json.load(open('gpurun_out/t8/r02_train_step_${N}gpu.json')); print(d['value'], d['n_gpus'], d['ms_per_step'], d['allreduce'], d['clocks'])"; tail -n 3 gpurun_out/t8/r02_train_step_${N}gpu.err
