#!/bin/bash
# Round-end evidence: (1) the bench command without a profiler, (2) its ncu launch list, (3) one ncu --set full capture
# of the dominant kernel from the same command.
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-reference-container"
$CMD > gpurun_out/final_plain.log 2>&1 || { tail -5 gpurun_out/final_plain.log; exit 1; }
tail -1 gpurun_out/final_plain.log | cut -c1-300
# one encode + decode is ~1.5 k launches with the dataflow kernel: skip the allocation pass and the warm-up step
ncu --metrics gpu__time_duration.sum --clock-control none -s 3200 -c 1700 --csv --log-file gpurun_out/r1_final_launches.csv $CMD > gpurun_out/final_ncu_list.log 2>&1
tail -1 gpurun_out/final_ncu_list.log | cut -c1-200
ncu --set full --clock-control none --import-source on -k regex:gemm_flow -s 150 -c 2 -o gpurun_out/r1_final_gemm_flow -f $CMD > gpurun_out/final_ncu_full.log 2>&1
tail -1 gpurun_out/final_ncu_full.log | cut -c1-200
