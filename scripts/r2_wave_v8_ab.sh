#!/bin/bash
# same-box A/B of the wave kernel: scripts/_ab/old.so (previous commit) against new.so, one 768x512 image, all topologies
mkdir -p gpurun_out
SO=learned-block-based-image-compression_b200/liblbic_b200.so
L=gpurun_out/r2_wave_v8_ab.log
: > $L
for v in old new old new; do
  cp scripts/_ab/$v.so $SO
  echo "== $v" >> $L
  LBIC_LAT_LANE_ONLY=1 timeout 300 python scripts/latency_topologies.py 2>> $L >> $L
done
for v in old new; do
  cp scripts/_ab/$v.so $SO
  echo "== $v, both containers" >> $L
  LBIC_LAT_CONFIGS=B8_lowrate timeout 300 python scripts/latency_topologies.py 2>> $L | grep reference >> $L
done
cp scripts/_ab/new.so $SO
LBIC_TRACE_CONFIG=B8_highrate bash scripts/r2_trace.sh > /dev/null 2>&1
cp gpurun_out/wave_trace_enc_summary.txt gpurun_out/r2_wave_trace_k3_enc_summary_v8.txt
cp gpurun_out/wave_trace_dec_summary.txt gpurun_out/r2_wave_trace_k3_dec_summary_v8.txt
cat $L
