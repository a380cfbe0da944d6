#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/$1; mkdir -p $O
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_train_ingest.py tests/test_gpu_parity.py -q -m gpu -x -k "similarity or specaug or snr or ingest or sharded_search_phases or mel_variants or encoder_variants or front_kernel or knn_many" > $O/memcheck.log 2>&1; echo "memcheck exit $?" | tee -a $O/summary.txt
grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" $O/memcheck.log | tail -8
