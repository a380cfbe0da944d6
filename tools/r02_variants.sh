#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/$1; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_train_ingest.py tests/test_gpu_cli.py -x -q -m gpu > $O/pytest_variants.log 2>&1; echo "variant tests exit $?" | tee -a $O/summary.txt
tail -n 30 $O/pytest_variants.log
