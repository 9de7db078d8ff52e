#!/bin/bash
# wave kernel for the KS3311 topologies: the wave equivalence cases, one-image latency, per-tile trace of B8_highrate
mkdir -p gpurun_out
L=gpurun_out/r2_k3wave.log
echo "== wave tests" > $L
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -k "wave_kernel" >> $L 2>&1
echo "exit $?" >> $L
timeout 300 python scripts/latency_topologies.py 2>> $L | grep -E "highrate.*lane" >> $L
LBIC_TRACE_CONFIG=B8_highrate bash scripts/r2_trace.sh > /dev/null 2>&1
tail -24 gpurun_out/wave_trace_enc_summary.txt >> $L
grep -E "^exit|^==|passed|failed|FAILED|highrate|^  [A-Z]|^step" $L
