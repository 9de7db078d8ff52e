#!/bin/bash
# wave kernel: weight boxes requested before a tile's dependency wait (LBIC_WAVE_WFIRST) x adaptive k-block groups
mkdir -p gpurun_out
L=gpurun_out/r2_wfirst_sweep2.log
: > $L
for c in "0 0" "0 12" "0 4" "1 0"; do
  set -- $c
  echo "== LBIC_WAVE_KB_ADAPT=$1 LBIC_WAVE_WFIRST=$2" >> $L
  LBIC_WAVE_KB_ADAPT=$1 LBIC_WAVE_WFIRST=$2 LBIC_LAT_CONFIGS=B8_lowrate,B8_highrate LBIC_LAT_LANE_ONLY=1 timeout 300 python scripts/latency_topologies.py 2>> $L >> $L
done
cat $L
