#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -m gpu -k "knn or sharded or search or query or planted or seq_score or roundtrip or e2e or end_to_end" 2>&1 | tail -5
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --steps 3 --warmup 2 --clips 50 --no-cpu > gpurun_out/bench_rerank.json 2> gpurun_out/bench_rerank.err
python - <<'PY'
import json
j=json.loads(open('gpurun_out/bench_rerank.json').read().strip().split('\n')[-1]); m=j['match']
print(m['value'], m['e2e']['value'], m['kernels_ms'], m['per_file_regime'])
PY
