#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu" > gpurun_out/full_check.log
timeout 1500 python -m pytest tests -m gpu -q >> gpurun_out/full_check.log 2>&1
echo "exit $?" >> gpurun_out/full_check.log
echo "== smoke" >> gpurun_out/full_check.log
timeout 300 python __graft_entry__.py smoke >> gpurun_out/full_check.log 2>&1
echo "exit $?" >> gpurun_out/full_check.log
echo "== bench default" >> gpurun_out/full_check.log
( time timeout 1200 python bench.py ) >> gpurun_out/full_check.log 2>&1
echo "exit $?" >> gpurun_out/full_check.log
echo "== bench reference arm" >> gpurun_out/full_check.log
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) >> gpurun_out/full_check.log 2>&1
grep -E "^exit|passed|failed|^==|Error|^real|smoke:" gpurun_out/full_check.log
