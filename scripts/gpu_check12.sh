#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/check12.log
echo "== pytest gpu" > $L
timeout 1500 python -m pytest tests -m gpu -q >> $L 2>&1
echo "exit $?" >> $L
echo "== smoke" >> $L
timeout 300 python __graft_entry__.py smoke >> $L 2>&1
echo "exit $?" >> $L
echo "== bench default" >> $L
( time timeout 1200 python bench.py --no-cpu-baseline ) >> $L 2>&1
echo "exit $?" >> $L
echo "== layer profile" >> $L
timeout 300 python scripts/layer_profile.py B8_lowrate 1024 >> $L 2>&1
grep -E "^exit|passed|failed|^==|Error|^real|smoke:" $L
