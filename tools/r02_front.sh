#!/bin/bash
# front_tc_kernel: parity tests first (hard timeouts: a hung cooperative kernel must not eat the call), then timing.
cd "$(dirname "$0")/.."
O=gpurun_out/r02b; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "front_kernel or encoder_bf16 or layer0 or fused_conv" > $O/pytest_front.log 2>&1; echo "front tests exit $?" | tee -a $O/summary.txt
tail -n 25 $O/pytest_front.log
timeout 300 python bench.py --clips 2000 --steps 3 --warmup 2 --no-match --no-cpu > $O/bench_front.log 2>&1; echo "bench exit $?" | tee -a $O/summary.txt
tail -n 1 $O/bench_front.log | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(j['value'], j['roofline']['kernels_ms_per_step'])"
PFANN_B200_NO_FRONT=1 timeout 300 python bench.py --clips 2000 --steps 3 --warmup 2 --no-match --no-cpu > $O/bench_nofront.log 2>&1
tail -n 1 $O/bench_nofront.log | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(j['value'], j['roofline']['kernels_ms_per_step'])"
