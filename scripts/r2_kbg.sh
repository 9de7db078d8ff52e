#!/bin/bash
# wave kernel: k-blocks per barrier round trip of the MMA issuer (LBIC_WAVE_KB_GROUP), one-image latency
mkdir -p gpurun_out
L=gpurun_out/r2_kbg.log
echo "== tests" > $L
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -k "wave_kernel" >> $L 2>&1
echo "exit $?" >> $L
for v in 1 2 3 4; do
  echo "== LBIC_WAVE_KB_GROUP=$v" >> $L
  LBIC_WAVE_KB_GROUP=$v timeout 300 python scripts/latency.py --quick 2>> $L | grep '"wave_kernel": true' | grep lane | cut -c1-200 >> $L
done
cat $L
