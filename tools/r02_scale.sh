#!/bin/bash
# multi-GPU: sharded == unsharded check, then the bench line at the box's GPU count
cd "$(dirname "$0")/.."
O=gpurun_out/$1; mkdir -p $O
N=$(nvidia-smi -L | wc -l)
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/dist_check.py > $O/dist_check_n$N.log 2>&1; echo "dist_check exit $?" | tee -a $O/summary.txt
tail -n 3 $O/dist_check_n$N.log
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 3 --warmup 3 > $O/bench_n$N.log 2>&1; echo "bench exit $?" | tee -a $O/summary.txt
tail -n 1 $O/bench_n$N.log > $O/bench_n$N.json
python -c "import sys,json; j=json.loads(open('$O/bench_n$N.json').read()); m=j['match']; print(j['value'], j['e2e']['value'], m['value'], m['e2e'], m['kernels_ms'], m['per_file_regime'])"
