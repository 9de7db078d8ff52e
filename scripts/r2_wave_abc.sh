#!/bin/bash
# same-box comparison of library builds scripts/_ab/{a,b,c,..}.so ($LBIC_AB_VARIANTS), one 768x512 image, lane container
mkdir -p gpurun_out
SO=learned-block-based-image-compression_b200/liblbic_b200.so
L=gpurun_out/r2_wave_abc.log
cp $SO gpurun_out/_keep.so
: > $L
for v in ${LBIC_AB_VARIANTS:-a b c a b c}; do
  cp scripts/_ab/$v.so $SO
  echo "== $v" >> $L
  LBIC_LAT_LANE_ONLY=1 timeout 300 python scripts/latency_topologies.py 2>> $L >> $L
done
cp gpurun_out/_keep.so $SO; rm gpurun_out/_keep.so
cat $L
