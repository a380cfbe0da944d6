#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/$1; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_parity.py -x -q -m gpu -k "end_to_end or mel" > $O/pytest_parity.log 2>&1; echo "parity tests exit $?" | tee -a $O/summary.txt
tail -n 25 $O/pytest_parity.log
timeout 600 compute-sanitizer --tool racecheck --racecheck-report all python -m pytest tests/test_gpu_parity.py -q -m gpu -k "front_kernel" > $O/racecheck_front.log 2>&1; echo "racecheck exit $?" | tee -a $O/summary.txt
grep -E "RACECHECK SUMMARY|passed|failed" $O/racecheck_front.log | tail -3
grep -A8 "hazard detected" $O/racecheck_front.log | grep -E "Thread|\.cu:" | sed 's/Thread ([0-9]*,0,0)/Thread T/; s/ at .* in / /' | sort | uniq -c | sort -rn | head -8
