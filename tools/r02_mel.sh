#!/bin/bash
# round 2: mel kernel v2 -- parity, then the bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -m gpu -k "mel or extract or e2e or end_to_end" 2>&1 | tail -8
timeout 600 python bench.py --steps 5 --warmup 3 --no-match --no-cpu > gpurun_out/bench_mel2.json 2> gpurun_out/bench_mel2.err
cut -c1-1500 gpurun_out/bench_mel2.json; tail -3 gpurun_out/bench_mel2.err
PFANN_B200_NO_OVERLAP=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-match --no-cpu > gpurun_out/bench_mel2_serial.json 2>> gpurun_out/bench_mel2.err
cut -c1-1500 gpurun_out/bench_mel2_serial.json
