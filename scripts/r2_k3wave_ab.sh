#!/bin/bash
# wave kernel A/B of one tuning hook (default LBIC_WAVE_WFIRST) off / on, all topologies, one 768x512 image; then the KS3311 trace
mkdir -p gpurun_out
K=${LBIC_AB_KNOB:-LBIC_WAVE_WFIRST}
L=gpurun_out/r2_k3wave_ab.log
echo "== wave tests" > $L
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -k "wave_kernel" >> $L 2>&1
echo "exit $?" >> $L
for a in 0 1 0 1; do
  echo "== $K=$a" >> $L
  env $K=$a timeout 300 python scripts/latency_topologies.py 2>> $L | grep -E "lane" >> $L
done
LBIC_TRACE_CONFIG=B8_highrate bash scripts/r2_trace.sh > /dev/null 2>&1
tail -24 gpurun_out/wave_trace_enc_summary.txt >> $L
grep -E "^exit|^==|passed|failed|FAILED|lane|^  [A-Z]|^step" $L
