cd /root/repo 2>/dev/null || cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out/v5
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/v5/launches.csv python bench.py --steps 1 --warmup 1 --clips 300 --chunk 4096 --db-rows 1000000 --queries 512 --no-cpu > gpurun_out/v5/ncu_list.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_ln_tc -s 0 -c 7 -o gpurun_out/v5/prof_convln python tools/conv_probe.py 140 > gpurun_out/v5/ncu_convln.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'l0_tc_kernel|mel_kernel|head_kernel|conv_gemm_tc' -s 0 -c 11 -o gpurun_out/v5/prof_other python tools/conv_probe.py 140 > gpurun_out/v5/ncu_other.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'knn_scan_tc|knn_select|rerank' -s 4 -c 6 -o gpurun_out/v5/prof_knn python bench.py --steps 1 --warmup 1 --clips 20 --db-rows 10000000 --queries 1024 --match-batch 1024 --no-cpu > gpurun_out/v5/ncu_knn.log 2>&1
ls -la gpurun_out/v5 | head -20
