#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r02d; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "front_kernel or encoder_bf16 or layer0 or fused_conv" > $O/pytest_front.log 2>&1; echo "front tests exit $?" | tee -a $O/summary.txt
tail -n 12 $O/pytest_front.log
timeout 300 python tools/front_probe.py > $O/front_probe.log 2>&1; tail -4 $O/front_probe.log
timeout 300 python bench.py --clips 2000 --steps 3 --warmup 2 --no-match --no-cpu > $O/bench_front.log 2>&1; echo "bench exit $?" | tee -a $O/summary.txt
tail -n 1 $O/bench_front.log | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(j['value'], j['roofline']['kernels_ms_per_step'])"
