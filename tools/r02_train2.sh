#!/bin/bash
# round 2: training step at 1 and 2 GPUs (gradient-sum check at 2), with the torch-CPU baseline at N=1
mkdir -p gpurun_out
timeout 400 python tools/train_bench.py --steps 20 --warmup 3 --profile --cpu-baseline 640 > gpurun_out/train_n1.json 2> gpurun_out/train_n1.err
cut -c1-900 gpurun_out/train_n1.json
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
    tools/train_bench.py --gpus 2 --steps 20 --warmup 3 --check > gpurun_out/train_n2.json 2> gpurun_out/train_n2.err
cut -c1-1200 gpurun_out/train_n2.json; tail -3 gpurun_out/train_n2.err
fi
