#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --images 64 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 9000 -c 500 --csv --log-file gpurun_out/launches_r1.csv $CMD > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 2200 -c 3 -o gpurun_out/prof_gemm_r1 $CMD > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
