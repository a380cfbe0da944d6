#!/bin/bash
# whole GPU suite + smoke on whatever GPUs the box has; then the N-GPU bench line
cd "$(dirname "$0")/.."
O=gpurun_out/$1; mkdir -p $O
timeout 1500 python -m pytest tests -q -m gpu -x > $O/pytest_all.log 2>&1; echo "pytest exit $?" | tee -a $O/summary.txt
tail -n 6 $O/pytest_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke exit $?" | tee -a $O/summary.txt; tail -n 2 $O/smoke.log
N=$(nvidia-smi -L | wc -l)
if [ "$N" -ge 2 ]; then
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 2 --warmup 3 --clips 1000 > $O/bench_n$N.log 2>&1; echo "bench exit $?" | tee -a $O/summary.txt
tail -n 1 $O/bench_n$N.log | python -c "import sys,json; j=json.loads(sys.stdin.read()); m=j['match']; print(j['value'], m['value'], m['e2e'], m['kernels_ms'])"
fi
