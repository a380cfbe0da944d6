#!/bin/bash
# compute-sanitizer racecheck over the kernels with new shared-memory hand-offs: mel_fast_kernel (transpose / spectrum /
# projection buffers, __syncwarp), rerank_kernel with several CTAs per file, head_kernel (4 groups, named barriers)
cd "$(dirname "$0")/.."
O=gpurun_out/$1; mkdir -p $O
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "mel_vs_golden or mel_pcm16 or query_edge_cases or models_of_different_configs" > $O/racecheck.log 2>&1; echo "racecheck exit $?" | tee -a $O/summary.txt
grep -E "RACECHECK SUMMARY|passed|failed|hazard" $O/racecheck.log | sort | uniq -c | tail -12
