#!/bin/bash
# Round-1 evidence run (writes gpurun_out/v5): full default bench line, ncu launch list, ncu --set full captures of the dominant kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/v5
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/v5/gpu.txt 2>&1
timeout 900 python bench.py > gpurun_out/v5/bench_n1.log 2>&1; tail -n 1 gpurun_out/v5/bench_n1.log > gpurun_out/v5/bench_n1.json
timeout 300 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/v5/bench_ref.log 2>&1; tail -n 1 gpurun_out/v5/bench_ref.log > gpurun_out/v5/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/v5/launches.csv python bench.py --steps 1 --warmup 1 --clips 300 --chunk 4096 --db-rows 1000000 --queries 512 --no-cpu > gpurun_out/v5/ncu_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_ln_tc -s 0 -c 7 -o gpurun_out/v5/prof_convln python tools/conv_probe.py 140 > gpurun_out/v5/ncu_convln.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'l0_tc_kernel|mel_kernel|head_kernel|conv_gemm_tc' -s 0 -c 11 -o gpurun_out/v5/prof_other python tools/conv_probe.py 140 > gpurun_out/v5/ncu_other.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'knn_scan_tc|knn_select|rerank' -s 4 -c 5 -o gpurun_out/v5/prof_knn python bench.py --steps 1 --warmup 1 --clips 20 --db-rows 10000000 --queries 512 --match-batch 512 --no-cpu > gpurun_out/v5/ncu_knn.log 2>&1
ls -la gpurun_out/v5
