#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_check4.log
echo "== new tests" > $L
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -k "band or eval_model or wave" >> $L 2>&1
echo "exit $?" >> $L
echo "== band bench 1 GPU (2048)" >> $L
timeout 600 python scripts/band_bench.py --size 2048 > gpurun_out/r2_band_1gpu_2048.json 2>> $L
echo "exit $?" >> $L
grep -E "^exit|passed|failed|^==|Error|^FAILED" $L; cat gpurun_out/r2_band_1gpu_2048.json
