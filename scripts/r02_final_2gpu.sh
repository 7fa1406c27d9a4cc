#!/bin/bash
# round 2, final revision, 2 GPUs: inference replicas and the distillation step with the NCCL arena all-reduce
mkdir -p gpurun_out/f2
O=gpurun_out/f2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > $O/r02_bench_2gpu.json 2> $O/r02_bench_2gpu.err; echo "bench 2gpu rc=$?"; python -c "import json; d=json.load(open('$O/r02_bench_2gpu.json')); print(d['value'], d['n_gpus'], d['e2e'], d['clocks'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --workload train-step --gpus 2 --steps 5 --warmup 2 > $O/r02_train_step_2gpu.json 2> $O/r02_train_step_2gpu.err; echo "train-step 2gpu rc=$?"; python -c "import json; d=json.load(open('$O/r02_train_step_2gpu.json')); print(d['value'], d['ms_per_step'], d['allreduce'], d['phases'])"; tail -n 3 $O/r02_train_step_2gpu.err
