#!/bin/bash
# First bring-up on the B200: SIMT twin end-to-end, then the tcgen05 core in its own processes.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
echo "== simt + integer kernels" > gpurun_out/check1.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "simt or rans or tables or layout or error" >> gpurun_out/check1.log 2>&1
echo "exit $?" >> gpurun_out/check1.log
echo "== tcgen05 gemm unit" >> gpurun_out/check1.log
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "gemm_core and tcgen05" >> gpurun_out/check1.log 2>&1
echo "exit $?" >> gpurun_out/check1.log
echo "== tcgen05 end to end" >> gpurun_out/check1.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "tcgen05 or roundtrip or invariance" >> gpurun_out/check1.log 2>&1
echo "exit $?" >> gpurun_out/check1.log
echo "== smoke" >> gpurun_out/check1.log
timeout 300 python __graft_entry__.py smoke >> gpurun_out/check1.log 2>&1
echo "exit $?" >> gpurun_out/check1.log
tail -5 gpurun_out/check1.log
