#!/bin/bash
# Which part of the epilogue sets the tile time?  Isolated layer launches (gemm_ws_kernel<PAIR>, 24576 rows) with parts of
# the epilogue switched off (LBIC_EPI_HACK: 1 no staging / stores, 2 no arithmetic, 4 no GDN side input, 8 no group loop);
# then the default bench with the same switches (dataflow kernel, TMA-store form).  Timing only: results are garbage.
mkdir -p gpurun_out
L=gpurun_out/r2_epi_hack.log
: > $L
for h in 0 1 2 3 7 8; do
  echo "== LBIC_EPI_HACK=$h (isolated launches)" >> $L
  LBIC_EPI_HACK=$h timeout 300 python scripts/epi_modes.py 24576 2>&1 | grep "pair=1" | grep -E "K=  768 C= 768|K=  576 C= 576|K= 1152 C= 960|K=  672 C= 672" >> $L
done
for h in 0 1 3 7 8; do
  echo "== LBIC_EPI_HACK=$h (default bench, 1024 images)" >> $L
  LBIC_EPI_HACK=$h timeout 600 python bench.py --images 1024 --steps 2 --warmup 2 --no-e2e --no-cpu-baseline --no-reference-container 2>> $L | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.1f enc %.1f dec %.1f sm_mhz %s' % (d['value'], d['encode_mpix_s'], d['decode_mpix_s'], d['clocks']['sm_mhz']))" >> $L
done
cat $L
