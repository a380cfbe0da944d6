#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r02c; mkdir -p $O
timeout 300 python tools/front_probe.py > $O/front_probe.log 2>&1; cat $O/front_probe.log | tail -5
timeout 600 ncu --set full --clock-control none --import-source on -k regex:front_tc -s 1 -c 1 -o $O/prof_front python tools/conv_probe.py 140 > $O/ncu_front.log 2>&1; tail -3 $O/ncu_front.log
ls -la $O
