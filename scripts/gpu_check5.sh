#!/bin/bash
mkdir -p gpurun_out
echo "== chain test" > gpurun_out/check5.log
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "chain_kernel" >> gpurun_out/check5.log 2>&1
echo "exit $?" >> gpurun_out/check5.log
echo "== pytest gpu" >> gpurun_out/check5.log
timeout 900 python -m pytest tests -m gpu -q -x >> gpurun_out/check5.log 2>&1
echo "exit $?" >> gpurun_out/check5.log
for n in 64 128 256; do
echo "== bench images=$n lanes=0" >> gpurun_out/check5.log
timeout 600 python bench.py --images $n --steps 2 --warmup 2 --no-cpu-baseline --no-e2e >> gpurun_out/check5.log 2>&1
echo "exit $?" >> gpurun_out/check5.log
done
grep -E "^exit|passed|failed|^==|Error" gpurun_out/check5.log
