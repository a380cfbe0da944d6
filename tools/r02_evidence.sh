#!/bin/bash
# Round-2 evidence run (writes gpurun_out/r02e): default bench line of both arms, ncu launch list, ncu --set full
# captures of the dominant kernels (front_tc, conv_ln_tc, knn_scan_tc incl. the per-file <32,0> instantiation, mel).
cd "$(dirname "$0")/.."
O=gpurun_out/r02e; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/gpu.txt 2>&1
timeout 900 python bench.py > $O/bench_n1.log 2>&1; echo "bench exit $?" | tee -a $O/summary.txt; tail -n 1 $O/bench_n1.log > $O/bench_n1.json
timeout 300 python bench.py --impl reference --steps 1 --warmup 1 > $O/bench_ref.log 2>&1; tail -n 1 $O/bench_ref.log > $O/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $O/launches.csv python bench.py --steps 1 --warmup 1 --clips 300 --chunk 4096 --db-rows 1000000 --queries 512 --no-cpu > $O/ncu_list.log 2>&1; echo "ncu list exit $?" | tee -a $O/summary.txt
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'front_tc|conv_ln_tc' -s 8 -c 7 -o $O/prof_conv python tools/conv_probe.py 140 > $O/ncu_conv.log 2>&1; echo "ncu conv exit $?" | tee -a $O/summary.txt
timeout 400 ncu --set full --clock-control none -k regex:'mel_kernel|head_kernel|conv_gemm_tc' -s 0 -c 10 -o $O/prof_other python tools/conv_probe.py 140 > $O/ncu_other.log 2>&1; echo "ncu other exit $?" | tee -a $O/summary.txt
timeout 500 ncu --set full --clock-control none -k regex:'knn_scan_tc|knn_select|rerank|knn_kth' -s 6 -c 10 -o $O/prof_knn python bench.py --steps 1 --warmup 1 --clips 20 --db-rows 10000000 --queries 1024 --match-batch 1024 --no-cpu > $O/ncu_knn.log 2>&1; echo "ncu knn exit $?" | tee -a $O/summary.txt
timeout 400 ncu --set full --clock-control none -k regex:'knn_scan_tc_kernel<32' -s 2 -c 3 -o $O/prof_knn32 python tools/knn_probe.py > $O/ncu_knn32.log 2>&1; echo "ncu knn32 exit $?" | tee -a $O/summary.txt
ls -la $O; cat $O/summary.txt
