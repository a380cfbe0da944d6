#!/bin/bash
# compute-sanitizer memcheck over the kernels added after the first memcheck record: mel_fast_kernel, head (4 groups),
# ln_apply_bf16x8, two-round select, union thresholds, the training forward/backward kernels, impulse responses
cd "$(dirname "$0")/.."
O=gpurun_out/$1; mkdir -p $O
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_train_ingest.py tests/test_gpu_parity.py -q -m gpu -x -k "training_step_gradients or training_step_end or impulse or mel_fast or mel_vs_golden or mel_pcm16 or sharded_search_phases or encoder_bf16_tensor_core or extract_pcm16_equals or knn_many" > $O/memcheck.log 2>&1; echo "memcheck exit $?" | tee -a $O/summary.txt
grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" $O/memcheck.log | tail -8
