#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/check13.log
echo "== pytest gpu (ws + rans + golden subset)" > $L
timeout 900 python -m pytest tests -m gpu -q -x -k "warp_specialised or golden or gemm_core or roundtrip or fixed_point" >> $L 2>&1
echo "exit $?" >> $L
echo "== epi modes" >> $L
timeout 300 python scripts/epi_modes.py >> $L 2>&1
echo "== layer profile" >> $L
timeout 300 python scripts/layer_profile.py B8_lowrate 1024 >> $L 2>&1
grep -E "^exit|passed|failed|^==|Error|GEMM total" $L
