#!/bin/bash
# Round-end evidence: (1) the bench command without a profiler, (2) its ncu launch list, (3) one ncu --set full capture
# of the dominant kernel from the same command.
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-reference-container"
$CMD > gpurun_out/final_plain.log 2>&1 || { tail -5 gpurun_out/final_plain.log; exit 1; }
tail -1 gpurun_out/final_plain.log | cut -c1-400
# launch list of one timed encode + decode (skip the warm-up: an encode is ~4.2 k launches, a decode ~2.9 k)
ncu --metrics gpu__time_duration.sum --clock-control none -s 7200 -c 7200 --csv --log-file gpurun_out/r1_final_launches.csv $CMD > gpurun_out/final_ncu_list.log 2>&1
tail -1 gpurun_out/final_ncu_list.log | cut -c1-200
