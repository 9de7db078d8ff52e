#!/bin/bash
# launch list of the default bench and ncu --set full of the dataflow kernel (LBIC_FLOW_COOP=0: Nsight Compute cannot
# replay a cooperative cluster launch; same kernel, plain launch as in round 1)
mkdir -p gpurun_out
export LBIC_FLOW_COOP=0
L=gpurun_out/r2_profile2.log
BENCH="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-reference-container"
echo "== plain bench" > $L
$BENCH > gpurun_out/r2_profile_plain.json 2>> $L && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 900 --csv --log-file gpurun_out/r2_launches_bench1024.csv $BENCH > gpurun_out/r2_ncu_launches.log 2>&1
echo "exit $?" >> $L
echo "== ncu full: gemm_flow" >> $L
$BENCH > /dev/null 2>> $L && \
ncu --set full --clock-control none --import-source on -k regex:gemm_flow -s 200 -c 2 -f -o gpurun_out/r2_flow $BENCH > gpurun_out/r2_ncu_flow.log 2>&1
echo "exit $?" >> $L
grep -E "^exit|^==" $L
