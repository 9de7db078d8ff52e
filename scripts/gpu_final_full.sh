#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-reference-container"
ncu --set full --clock-control none --import-source on -k regex:gemm_ws -s 9000 -c 3 -o gpurun_out/r1_final_gemm_ws -f $CMD > gpurun_out/final_ncu_full.log 2>&1
tail -2 gpurun_out/final_ncu_full.log | cut -c1-200
