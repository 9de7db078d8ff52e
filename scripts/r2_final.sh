#!/bin/bash
# final state of round 2: GPU test suite, smoke, default bench, reference arm, one-image / small-batch latency table, then
# (only after the bench exited 0 without ncu) the ncu launch list of the same command
mkdir -p gpurun_out
L=gpurun_out/r2_final.log
echo "== pytest gpu" > $L
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 >> $L 2>&1
echo "exit $?" >> $L
echo "== smoke" >> $L
timeout 300 python __graft_entry__.py smoke >> $L 2>&1
echo "exit $?" >> $L
echo "== bench" >> $L
timeout 900 python bench.py > gpurun_out/r2_bench_final.json 2>> $L
echo "exit $?" >> $L
echo "== reference arm" >> $L
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2>> $L
echo "exit $?" >> $L
echo "== latency" >> $L
timeout 900 python scripts/latency.py > gpurun_out/r2_latency_final.jsonl 2>> $L
echo "exit $?" >> $L
export LBIC_FLOW_COOP=0
BENCH="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-reference-container"
echo "== launch list" >> $L
$BENCH > /dev/null 2>> $L && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 900 --csv --log-file gpurun_out/r2_final_launches_bench1024.csv $BENCH > gpurun_out/r2_final_ncu_launches.log 2>&1
echo "exit $?" >> $L
grep -E "^exit|^==|passed|failed|smoke:" $L
