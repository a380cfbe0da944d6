#!/bin/bash
# sharded-search rework: single-GPU parity of the phases, kNN tests (multi-group launches), then the 2-GPU check + bench
cd "$(dirname "$0")/.."
O=gpurun_out/$1; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -m gpu -k "knn or sharded or search or query or seq_score or planted" > $O/pytest_knn.log 2>&1; echo "knn tests exit $?" | tee -a $O/summary.txt
tail -n 15 $O/pytest_knn.log
N=$(nvidia-smi -L | wc -l)
if [ "$N" -ge 2 ]; then
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/dist_check.py > $O/dist_check_n$N.log 2>&1; echo "dist_check exit $?" | tee -a $O/summary.txt
tail -n 4 $O/dist_check_n$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 2 --warmup 3 > $O/bench_n$N.log 2>&1; echo "bench exit $?" | tee -a $O/summary.txt
tail -n 1 $O/bench_n$N.log | python -c "import sys,json; j=json.loads(sys.stdin.read()); m=j['match']; print(j['value'], m['value'], m['e2e'], m['kernels_ms'], m['per_file_regime'])"
else
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu > $O/bench_n1.log 2>&1; echo "bench exit $?" | tee -a $O/summary.txt
tail -n 1 $O/bench_n1.log | python -c "import sys,json; j=json.loads(sys.stdin.read()); m=j['match']; print(j['value'], m['value'], m['e2e'], m['kernels_ms'], m['per_file_regime'])"
fi
