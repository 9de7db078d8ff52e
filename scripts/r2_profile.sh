#!/bin/bash
# round 2 profiles: launch list of the default bench, ncu --set full of the dataflow kernel and of the wave kernel
mkdir -p gpurun_out
L=gpurun_out/r2_profile.log
BENCH="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-reference-container"
echo "== plain bench" > $L
$BENCH > gpurun_out/r2_profile_plain.json 2>> $L && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 900 --csv --log-file gpurun_out/r2_launches_bench1024.csv $BENCH > gpurun_out/r2_ncu_launches.log 2>&1
echo "exit $?" >> $L
echo "== ncu full: gemm_flow" >> $L
$BENCH > /dev/null 2>> $L && \
ncu --set full --clock-control none --import-source on -k regex:gemm_flow -s 200 -c 2 -f -o gpurun_out/r2_flow $BENCH > gpurun_out/r2_ncu_flow.log 2>&1
echo "exit $?" >> $L
echo "== ncu full: gemm_wave" >> $L
python scripts/one_image.py >> $L 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_wave -s 2 -c 1 -f -o gpurun_out/r2_wave python scripts/one_image.py > gpurun_out/r2_ncu_wave.log 2>&1
echo "exit $?" >> $L
ls -la gpurun_out/*.ncu-rep >> $L 2>&1
grep -E "^exit|^==" $L
