#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/check16.log
echo "== pytest gpu subset" > $L
timeout 900 python -m pytest tests -m gpu -q -x -k "warp_specialised or golden or gemm_core or roundtrip or fixed_point or chain" >> $L 2>&1
echo "exit $?" >> $L
echo "== stages exp" >> $L
timeout 300 python scripts/stages_exp.py >> $L 2>&1
echo "== layer profile" >> $L
timeout 300 python scripts/layer_profile.py B8_lowrate 1024 >> $L 2>&1
grep -E "^exit|passed|failed|^==|Error|GEMM total|assert" $L
