#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu" > gpurun_out/check2.log
timeout 900 python -m pytest tests -m gpu -q -x >> gpurun_out/check2.log 2>&1
echo "exit $?" >> gpurun_out/check2.log
for n in 8 64; do
  echo "== bench images=$n lanes=0" >> gpurun_out/check2.log
  timeout 900 python bench.py --images $n --steps 2 --warmup 3 --cpu-blocks 192 >> gpurun_out/check2.log 2>&1
  echo "exit $?" >> gpurun_out/check2.log
done
echo "== bench images=64 lanes=1" >> gpurun_out/check2.log
timeout 900 python bench.py --images 64 --steps 1 --warmup 1 --lanes 1 --no-cpu-baseline >> gpurun_out/check2.log 2>&1
echo "exit $?" >> gpurun_out/check2.log
echo "== bench simt images=8" >> gpurun_out/check2.log
timeout 900 python bench.py --images 8 --steps 1 --warmup 1 --core simt --no-cpu-baseline --no-e2e >> gpurun_out/check2.log 2>&1
echo "exit $?" >> gpurun_out/check2.log
grep -E "^exit|passed|failed" gpurun_out/check2.log
