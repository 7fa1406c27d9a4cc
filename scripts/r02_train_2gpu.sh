#!/bin/bash
# 2-GPU distillation step: the NCCL all-reduce of the gradient arena overlapped with the (graph-segmented) backward
mkdir -p gpurun_out/d18
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --workload train-step --gpus 2 --steps 5 --warmup 2 > gpurun_out/d18/train_step_2gpu.json 2> gpurun_out/d18/train_step_2gpu.err; echo "train-step 2gpu rc=$?"; cat gpurun_out/d18/train_step_2gpu.json; tail -5 gpurun_out/d18/train_step_2gpu.err
