#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/$1; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "variants or mel or encoder_fp32" > $O/pytest_variants.log 2>&1; echo "variant tests exit $?" | tee -a $O/summary.txt
tail -n 30 $O/pytest_variants.log
