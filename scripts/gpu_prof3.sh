#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --images 1024 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/prof3_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 5000 -c 1200 --csv --log-file gpurun_out/launches_r1_ws.csv $CMD > gpurun_out/ncu_launch3.log 2>&1
cat gpurun_out/prof3_plain.log | tail -1
