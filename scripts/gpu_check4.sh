#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu" > gpurun_out/check4.log
timeout 1200 python -m pytest tests -m gpu -q -x >> gpurun_out/check4.log 2>&1
echo "exit $?" >> gpurun_out/check4.log
echo "== sweep" >> gpurun_out/check4.log
timeout 600 python scripts/gemm_sweep.py >> gpurun_out/check4.log 2>&1
for n in 64 128 256; do
echo "== bench images=$n lanes=0" >> gpurun_out/check4.log
timeout 900 python bench.py --images $n --steps 2 --warmup 2 --no-cpu-baseline --no-e2e >> gpurun_out/check4.log 2>&1
echo "exit $?" >> gpurun_out/check4.log
done
grep -E "^exit|passed|failed|^==" gpurun_out/check4.log
