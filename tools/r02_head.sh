#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "encoder or model or extract or head or layer" 2>&1 | tail -8
timeout 600 python bench.py --steps 5 --warmup 3 --no-match --no-cpu > gpurun_out/bench_head.json 2> gpurun_out/bench_head.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_head.json').read().strip().split('\n')[-1])
print(d['value'], d['ms_per_step'], d['roofline'].get('kernels_ms_per_step'))
PY
