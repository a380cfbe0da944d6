#!/bin/bash
# round 2: training step -- tests, timing, per-kernel launch list
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_train_ingest.py -x -q -m gpu -k "training or impulse" 2>&1 | tail -5
timeout 200 python tools/train_bench.py --steps 20 --warmup 3 --profile --cpu-baseline 640 > gpurun_out/train_n1.json 2> gpurun_out/train_n1.err
tail -c 3000 gpurun_out/train_n1.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/train_launches.csv \
    python tools/train_bench.py --steps 1 --warmup 1 > gpurun_out/train_ncu.log 2>&1
tail -2 gpurun_out/train_ncu.log | cut -c1-300
