#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/n4; mkdir -p $O
N=$(nvidia-smi -L | wc -l)
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/dist_check.py > $O/dist_check_n$N.log 2>&1; echo "dist_check exit $?"
tail -n 2 $O/dist_check_n$N.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 2 --warmup 3 --clips 1000 --no-cpu > $O/bench_n$N.log 2>&1; echo "bench exit $?"
tail -n 1 $O/bench_n$N.log > $O/bench_n$N.json
python -c "import sys,json; j=json.loads(open('$O/bench_n$N.json').read()); m=j['match']; print(j['value'], m['value'], m['e2e']['value'], m['kernels_ms'])"
