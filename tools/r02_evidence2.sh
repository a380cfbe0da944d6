#!/bin/bash
# second evidence pass: front_tc_kernel on a full 4096-segment chunk; the filtered kNN scans (256 queries per pass and the
# per-file 19-query instantiation <32,0>), select and rerank -- template arguments matched on demangled names
cd "$(dirname "$0")/.."
O=gpurun_out/r02e; mkdir -p $O
timeout 400 ncu --set full --clock-control none --import-source on -k regex:front_tc -s 3 -c 1 -o $O/prof_front python tools/conv_probe.py 140 > $O/ncu_front.log 2>&1; echo "ncu front exit $?" | tee -a $O/summary2.txt
timeout 600 ncu --set full --clock-control none --kernel-name-base demangled -k regex:'knn_scan_tc_kernel<256, 0>|knn_scan_tc_kernel<32, 0>|knn_select_kernel|rerank_kernel|merge_keys|combine_best' -s 4 -c 12 -o $O/prof_knn2 python bench.py --steps 1 --warmup 1 --clips 20 --db-rows 10000000 --queries 1024 --match-batch 1024 --no-cpu > $O/ncu_knn2.log 2>&1; echo "ncu knn2 exit $?" | tee -a $O/summary2.txt
ls -la $O | grep prof_; cat $O/summary2.txt
