#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/check15.log
echo "== pytest gpu" > $L
timeout 1500 python -m pytest tests -m gpu -q -x -k "thread_per_stream or roundtrip or corrupt or batch_invariance" >> $L 2>&1
echo "exit $?" >> $L
echo "== bench default" >> $L
( time timeout 1200 python bench.py --no-cpu-baseline ) >> $L 2>&1
echo "exit $?" >> $L
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -k regex:rans_dec -s 100 -c 10 --csv --log-file gpurun_out/dec_step_thread.csv python scripts/decode_launches.py 1024 0 > gpurun_out/ncu_launch7.log 2>&1
grep -E "^exit|passed|failed|^==|Error|^real|assert" $L | head -30
grep rans_dec gpurun_out/dec_step_thread.csv | cut -d, -f5,12- | head -5
