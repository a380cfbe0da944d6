#!/bin/bash
# First-contact GPU run: every test file in its own process with a hard timeout (a hung kernel must not take
# the whole call down), logs under gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { # name timeout cmd...
  local name=$1 to=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout "$to" "$@" > "gpurun_out/$name.log" 2>&1
  echo "exit $? : $name" | tee -a gpurun_out/summary.txt
  tail -n 15 "gpurun_out/$name.log"
}
: > gpurun_out/summary.txt
for t in "$@"; do
  case $t in
    mel)     run t_mel 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "mel" ;;
    enc32)   run t_enc32 400 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "encoder_fp32" ;;
    encbf)   run t_encbf 400 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "encoder_bf16 or extract" ;;
    knn)     run t_knn 400 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "knn" ;;
    rerank)  run t_rerank 400 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "seq_score or database_query or query_edge" ;;
    smoke)   run t_smoke 300 python __graft_entry__.py smoke ;;
    bench_small) run bench_small 600 python bench.py --clips 500 --db-rows 1000000 --queries 1000 --steps 2 --warmup 1 ;;
    bench)   run bench 1200 python bench.py ;;
    sweep)   for c in ${SWEEP:-2048 4096 8192}; do run sweep_$c 300 python bench.py --clips 2000 --chunk $c --steps 2 --warmup 1 --no-match --no-cpu; done ;;
    cli)     run t_cli 300 python -m pytest tests/test_gpu_cli.py -q -m gpu ;;
    ncu_l0)  run ncu_l0 600 ncu --set full --clock-control none --import-source on -k regex:'l0_tc_kernel|mel_kernel|head_kernel|l0_stats' -c 4 -o gpurun_out/prof_l0 python bench.py --steps 1 --warmup 1 --clips 300 --chunk 4096 --no-match --no-cpu ;;
    full)    run t_full 600 python -m pytest tests/test_gpu_fullsize.py -q -m gpu ;;
    probe)   run knn_probe 600 python tools/knn_probe.py ;;
    convprobe) run conv_probe 300 python tools/conv_probe.py 700 4096 ;;
    lnprobe) run ln_probe 300 python tools/ln_probe.py ;;
    all)     run t_all 900 python -m pytest tests -q -m gpu ;;
    ncu_list) run ncu_list 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --clips 300 --chunk 4096 --db-rows 1000000 --queries 100 --no-cpu ;;
    ncu_conv) run ncu_conv 900 ncu --set full --clock-control none --import-source on -k regex:'conv_ln_tc|conv_gemm_tc' -s 0 -c 15 -o gpurun_out/prof_conv python bench.py --steps 1 --warmup 1 --clips 300 --chunk 4096 --no-match --no-cpu ;;
    ncu_knn)  run ncu_knn 900 ncu --set full --clock-control none --import-source on -k regex:'knn_scan_tc|knn_select|rerank' -s 4 -c 6 -o gpurun_out/prof_knn python bench.py --steps 1 --warmup 1 --clips 20 --db-rows 10000000 --queries 1024 --match-batch 1024 --no-cpu ;;
  esac
done
cat gpurun_out/summary.txt
