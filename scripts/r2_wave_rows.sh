#!/bin/bash
# up to how many rows per step does the wave kernel (v7) beat the per-layer / single-CTA dataflow path? (LBIC_WAVE_MAX_ROWS)
mkdir -p gpurun_out
L=gpurun_out/r2_wave_rows.log
: > $L
run() {
  echo "== wave_max_rows=$1 wave_dec_max_rows=$2 images=$3" >> $L
  LBIC_WAVE_MAX_ROWS=$1 LBIC_WAVE_DEC_MAX_ROWS=$2 timeout 600 python bench.py --images $3 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-reference-container 2>> $L | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.1f enc %.1f dec %.1f identical %s' % (d['value'], d['encode_mpix_s'], d['decode_mpix_s'], d['enc_dec_identical']))" >> $L
}
for n in 24 32 48 64; do
  run 1536 64 $n
  run 2304 64 $n
  run 3072 64 $n
  run 1536 256 $n
  run 1536 1024 $n
done
cat $L
