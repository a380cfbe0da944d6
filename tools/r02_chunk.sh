#!/bin/bash
mkdir -p gpurun_out
for c in 16384 32768 65536; do
timeout 400 python bench.py --steps 3 --warmup 3 --no-match --no-cpu --chunk $c > gpurun_out/bench_chunk$c.json 2> gpurun_out/bench_chunk$c.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_chunk$c.json').read().strip().split('\n')[-1])
print($c, d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks']['sm_mhz'])
PY
done
