#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r02e; mkdir -p $O
B="python bench.py --steps 1 --warmup 1 --clips 20 --db-rows 10000000 --queries 1024 --match-batch 1024 --no-cpu"
timeout 300 ncu --set full --clock-control none -k regex:'knn_scan_tc_kernel' -s 20 -c 2 -o $O/prof_knn256 $B > $O/ncu_knn256.log 2>&1; echo "ncu knn256 exit $?" | tee -a $O/summary3.txt
timeout 300 ncu --set full --clock-control none -k regex:'knn_scan_tc_kernel' -s 116 -c 4 -o $O/prof_knn32 $B > $O/ncu_knn32.log 2>&1; echo "ncu knn32 exit $?" | tee -a $O/summary3.txt
timeout 300 ncu --set full --clock-control none -k regex:'rerank_kernel|knn_select_kernel|merge_keys_kernel' -s 0 -c 6 -o $O/prof_rerank $B > $O/ncu_rerank.log 2>&1; echo "ncu rerank exit $?" | tee -a $O/summary3.txt
