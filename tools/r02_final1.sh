#!/bin/bash
# round 2, after the mel / head / LN / select changes: whole GPU suite, smoke, both bench arms at N=1
cd "$(dirname "$0")/.."
O=gpurun_out/final1; mkdir -p $O
timeout 1500 python -m pytest tests -q -m gpu -x > $O/pytest_all.log 2>&1; echo "pytest exit $?" | tee -a $O/summary.txt
tail -n 4 $O/pytest_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke exit $?" | tee -a $O/summary.txt; tail -n 2 $O/smoke.log
timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench exit $?" | tee -a $O/summary.txt
python - <<'PY'
import json
j=json.loads(open('gpurun_out/final1/bench_n1.json').read().strip().split('\n')[-1]); m=j['match']
print(j['value'], j['e2e']['value'], j['ms_per_step'], j['roofline']['frac'], j['clocks'])
print(m['value'], m['e2e'], m.get('kernels_ms'))
print(j['cpu_baseline'])
PY
timeout 600 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref exit $?" | tee -a $O/summary.txt
cut -c1-600 $O/bench_ref.json
