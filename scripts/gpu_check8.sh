#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu" > gpurun_out/check8.log
timeout 1500 python -m pytest tests -m gpu -q >> gpurun_out/check8.log 2>&1
echo "exit $?" >> gpurun_out/check8.log
for lanes in 0 1; do
echo "== bench images=256 lanes=$lanes" >> gpurun_out/check8.log
timeout 900 python bench.py --images 256 --steps 2 --warmup 2 --lanes $lanes --no-cpu-baseline >> gpurun_out/check8.log 2>&1
echo "exit $?" >> gpurun_out/check8.log
done
grep -E "^exit|passed|failed|^==|Error" gpurun_out/check8.log
cat gpurun_out/fixed_point_*.json gpurun_out/full_size_parity.json
