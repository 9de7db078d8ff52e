#!/bin/bash
# wave kernel with M = 64 MMAs for steps of at most 64 rows: equivalence tests, one-image latency with and without
mkdir -p gpurun_out
L=gpurun_out/r2_m64.log
echo "== tests" > $L
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -k "wave_kernel or batch_invariance or eval_model or image_codec or u8_image" >> $L 2>&1
echo "exit $?" >> $L
for v in 0 1; do
  echo "== LBIC_WAVE_M64=$v" >> $L
  LBIC_WAVE_M64=$v timeout 300 python scripts/latency.py --quick 2>> $L | grep '"wave_kernel": true' | cut -c1-200 >> $L
done
cat $L
