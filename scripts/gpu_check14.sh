#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/check14.log
echo "== pytest gpu" > $L
timeout 1500 python -m pytest tests -m gpu -q -x >> $L 2>&1
echo "exit $?" >> $L
echo "== bench default" >> $L
( time timeout 1200 python bench.py --no-cpu-baseline ) >> $L 2>&1
echo "exit $?" >> $L
grep -E "^exit|passed|failed|^==|Error|^real|assert" $L | head -30
