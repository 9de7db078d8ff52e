#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu" > gpurun_out/check11.log
timeout 1500 python -m pytest tests -m gpu -q >> gpurun_out/check11.log 2>&1
echo "exit $?" >> gpurun_out/check11.log
timeout 600 python scripts/gemm_sweep2.py >> gpurun_out/check11.log 2>&1
timeout 900 python bench.py --no-cpu-baseline --no-e2e --no-reference-container >> gpurun_out/check11.log 2>&1
grep -E "^exit|passed|failed|^==|Error|^ws=1" gpurun_out/check11.log
