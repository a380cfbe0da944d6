#!/bin/bash
# Round-2 evidence refresh after the mel v2 / head / tail-LayerNorm / select changes: launch list of the bench command,
# ncu --set full of the changed kernels
cd "$(dirname "$0")/.."
O=gpurun_out/r02e4; mkdir -p $O
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $O/launches.csv python bench.py --steps 1 --warmup 1 --clips 300 --chunk 4096 --db-rows 1000000 --queries 512 --no-cpu > $O/ncu_list.log 2>&1; echo "ncu list exit $?" | tee -a $O/summary.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'mel_fast_kernel|head_kernel|ln_apply_bf16x8|conv_gemm_tc' -s 0 -c 12 -o $O/prof_other python tools/conv_probe.py 140 > $O/ncu_other.log 2>&1; echo "ncu other exit $?" | tee -a $O/summary.txt
B="python bench.py --steps 1 --warmup 1 --clips 20 --db-rows 10000000 --queries 1024 --match-batch 1024 --no-cpu"
timeout 300 ncu --set full --clock-control none -k regex:'rerank_kernel|knn_select_kernel|merge_keys_kernel' -s 0 -c 6 -o $O/prof_rerank $B > $O/ncu_rerank.log 2>&1; echo "ncu rerank exit $?" | tee -a $O/summary.txt
ls -la $O; cat $O/summary.txt
