#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_epi_hack2.log
: > $L
for h in 0 16 1; do
  echo "== LBIC_EPI_HACK=$h (isolated launches, copy-loop form)" >> $L
  LBIC_EPI_HACK=$h timeout 300 python scripts/epi_modes.py 24576 2>&1 | grep "pair=1" | grep -E "K=  768 C= 768|K=  576 C= 576" | grep -E "PREGDN|GDN|LRELU|RAW" >> $L
done
for h in 0 16; do
 for t in 0 1; do
  echo "== LBIC_EPI_HACK=$h LBIC_TMA_STORE=$t (default bench, 1024 images)" >> $L
  LBIC_TMA_STORE=$t LBIC_EPI_HACK=$h timeout 600 python bench.py --images 1024 --steps 2 --warmup 2 --no-e2e --no-cpu-baseline --no-reference-container 2>> $L | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.1f enc %.1f dec %.1f sm_mhz %s' % (d['value'], d['encode_mpix_s'], d['decode_mpix_s'], d['clocks']['sm_mhz']))" >> $L
 done
done
cat $L
