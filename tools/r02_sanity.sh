#!/bin/bash
# Round-2 first GPU call: whole GPU suite, the default bench line of both arms, racecheck of the fused conv+LN test.
cd "$(dirname "$0")/.."
O=gpurun_out/r02a; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/gpu.txt 2>&1
timeout 900 python -m pytest tests -q -m gpu -x > $O/pytest.log 2>&1; echo "pytest exit $?" | tee -a $O/summary.txt
timeout 900 python bench.py --steps 3 --warmup 3 > $O/bench_n1.log 2>&1; echo "bench exit $?" | tee -a $O/summary.txt; tail -n 1 $O/bench_n1.log > $O/bench_n1.json
timeout 300 python bench.py --impl reference --steps 1 --warmup 1 > $O/bench_ref.log 2>&1; tail -n 1 $O/bench_ref.log > $O/bench_ref.json
timeout 600 compute-sanitizer --tool racecheck --racecheck-report all python -m pytest tests/test_gpu_parity.py -q -m gpu -k "fused_conv_layernorm or knn_exact" > $O/racecheck.log 2>&1; echo "racecheck exit $?" | tee -a $O/summary.txt
tail -n 5 $O/pytest.log; tail -c 1500 $O/bench_n1.json; tail -n 8 $O/racecheck.log
