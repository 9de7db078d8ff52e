#!/bin/bash
# wave kernel for the KS3311 topologies: the wave equivalence cases, then the whole GPU suite
mkdir -p gpurun_out
L=gpurun_out/r2_k3wave.log
echo "== wave tests" > $L
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -k "wave_kernel" >> $L 2>&1
echo "exit $?" >> $L
echo "== pytest gpu (all)" >> $L
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 >> $L 2>&1
echo "exit $?" >> $L
grep -E "^exit|^==|passed|failed|FAILED" $L
