#!/bin/bash
# round 2: training step on 2 GPUs, with the gradient-sum check against one GPU on the global batch
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
    tools/train_bench.py --gpus 2 --steps 20 --warmup 3 --check > gpurun_out/train_n2.json 2> gpurun_out/train_n2.err
cut -c1-1200 gpurun_out/train_n2.json; tail -3 gpurun_out/train_n2.err
