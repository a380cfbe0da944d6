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
    all)     run t_all 900 python -m pytest tests -q -m gpu ;;
  esac
done
cat gpurun_out/summary.txt
