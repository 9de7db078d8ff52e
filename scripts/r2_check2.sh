#!/bin/bash
# round 2, GPU call 2: the wave kernel -- equivalence tests first (bounded), then latency, then the rest of the suite
mkdir -p gpurun_out
L=gpurun_out/r2_check2.log
echo "== wave tests" > $L
timeout 600 python -m pytest tests -m gpu -q -x --timeout 300 -k "wave or u8 or banded or corrupt" >> $L 2>&1
rc=$?
echo "exit $rc" >> $L
if [ $rc -eq 0 ]; then
  echo "== latency" >> $L
  timeout 900 python scripts/latency.py > gpurun_out/r2_latency.jsonl 2>> $L
  echo "exit $?" >> $L
fi
echo "== pytest gpu (all)" >> $L
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 >> $L 2>&1
echo "exit $?" >> $L
grep -E "^exit|passed|failed|^==|Error|^FAILED" $L
