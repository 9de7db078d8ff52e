#!/bin/bash
# same-box A/B of two builds of the library (scripts/_ab/old.so, new.so): wave tests on new, one-image latency of the KS3311 topologies
mkdir -p gpurun_out
SO=learned-block-based-image-compression_b200/liblbic_b200.so
L=gpurun_out/r2_wave_ab_quick.log
cp scripts/_ab/new.so $SO
timeout 600 python -m pytest tests -m gpu -q -x --timeout 600 -k "wave_kernel" 2>&1 | tail -2 > $L
for v in old new old new; do
  cp scripts/_ab/$v.so $SO
  echo "== $v" >> $L
  LBIC_LAT_LANE_ONLY=1 LBIC_LAT_CONFIGS=${LBIC_LAT_CONFIGS:-B8_highrate,B4_highrate} timeout 300 python scripts/latency_topologies.py 2>> $L >> $L
done
cp scripts/_ab/new.so $SO
LBIC_TRACE_CONFIG=B8_highrate bash scripts/r2_trace.sh > /dev/null 2>&1
grep -E "^step|^  [A-Z]" gpurun_out/wave_trace_enc_summary.txt | head -23 >> $L
cat $L
