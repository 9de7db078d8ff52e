#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu" > gpurun_out/check3.log
timeout 1200 python -m pytest tests -m gpu -q >> gpurun_out/check3.log 2>&1
echo "exit $?" >> gpurun_out/check3.log
echo "== bench images=64 lanes=0" >> gpurun_out/check3.log
timeout 900 python bench.py --images 64 --steps 2 --warmup 3 --no-cpu-baseline >> gpurun_out/check3.log 2>&1
echo "exit $?" >> gpurun_out/check3.log
echo "== bench images=128 lanes=0" >> gpurun_out/check3.log
timeout 900 python bench.py --images 128 --steps 2 --warmup 2 --no-cpu-baseline --no-e2e >> gpurun_out/check3.log 2>&1
echo "exit $?" >> gpurun_out/check3.log
grep -E "^exit|passed|failed|^==" gpurun_out/check3.log
