#!/bin/bash
mkdir -p gpurun_out
python scripts/one_gemm.py 24576 768 768 3 1 > gpurun_out/one_gemm.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:gemm_ws -s 4 -c 1 -o gpurun_out/quant_ws -f python scripts/one_gemm.py 24576 768 768 3 1 > gpurun_out/ncu_quant.log 2>&1
tail -3 gpurun_out/ncu_quant.log
