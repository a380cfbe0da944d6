#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/$1; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_cli.py -x -q -m gpu -k "extract or front or roundtrip or end_to_end" > $O/pytest_extract.log 2>&1; echo "extract tests exit $?" | tee -a $O/summary.txt
tail -n 6 $O/pytest_extract.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-match --no-cpu > $O/bench_overlap.log 2>&1; echo "bench exit $?" | tee -a $O/summary.txt
tail -n 1 $O/bench_overlap.log | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('overlap', j['value'], j['e2e']['value'], j['ms_per_step'], j['roofline']['kernels_ms_per_step'])"
PFANN_B200_NO_OVERLAP=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-match --no-cpu > $O/bench_nooverlap.log 2>&1
tail -n 1 $O/bench_nooverlap.log | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('serial ', j['value'], j['e2e']['value'], j['ms_per_step'], j['roofline']['kernels_ms_per_step'])"
