#!/bin/bash
# N-GPU match leg with different per-shard sample sizes (union-of-samples thresholds)
cd "$(dirname "$0")/.."
O=gpurun_out/$1; mkdir -p $O
N=$(nvidia-smi -L | wc -l)
if [ "$N" -eq 1 ]; then
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -m gpu -k "knn or sharded or search or query or planted" 2>&1 | tail -40
exit 0
fi
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/dist_check.py > $O/dist_check_n$N.log 2>&1; echo "dist_check exit $?" | tee -a $O/summary.txt
tail -n 2 $O/dist_check_n$N.log
for sc in 1.0 0.5; do
PFANN_B200_SAMPLE_SCALE=$sc timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 2 --warmup 2 --clips 50 --no-cpu > $O/bench_n${N}_$sc.log 2>&1; echo "bench $sc exit $?" | tee -a $O/summary.txt
tail -n 1 $O/bench_n${N}_$sc.log | python -c "import sys,json; j=json.loads(sys.stdin.read()); m=j['match']; print('$sc', m['value'], m['e2e']['value'], m['ms_per_step'], m['kernels_ms'])"
done
