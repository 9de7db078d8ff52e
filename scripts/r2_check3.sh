#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_check3.log
echo "== wave tests" > $L
timeout 600 python -m pytest tests -m gpu -q -x --timeout 300 -k "wave or u8 or banded or rans" >> $L 2>&1
rc=$?
echo "exit $rc" >> $L
if [ $rc -eq 0 ]; then
  echo "== latency" >> $L
  timeout 900 python scripts/latency.py > gpurun_out/r2_latency.jsonl 2>> $L
  echo "exit $?" >> $L
  bash scripts/r2_trace.sh >> $L 2>&1
fi
grep -E "^exit|passed|failed|^==|Error|^FAILED" $L
