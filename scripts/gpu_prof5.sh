#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --images 1024 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-reference-container"
$CMD > gpurun_out/prof5_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 5000 -c 800 --csv --log-file gpurun_out/launches_r1_ws2.csv $CMD > gpurun_out/ncu_launch5.log 2>&1
tail -1 gpurun_out/prof5_plain.log | cut -c1-300
