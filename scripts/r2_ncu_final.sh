#!/bin/bash
# ncu --set full of the dataflow kernel at the peak of the wavefront (launches 105-107 of the run), final state of round 2
mkdir -p gpurun_out
export LBIC_FLOW_COOP=0
BENCH="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-reference-container"
$BENCH > /dev/null 2> gpurun_out/r2_ncu_final.log && \
ncu --set full --clock-control none --import-source on -k regex:gemm_flow -s 105 -c 3 -f -o gpurun_out/r2_final_flow_peak $BENCH > gpurun_out/r2_final_ncu_flow_peak.log 2>&1
echo "exit $?"
ls -la gpurun_out/r2_final_flow_peak.ncu-rep
