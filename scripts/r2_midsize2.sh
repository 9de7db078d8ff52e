#!/bin/bash
# threshold between the single-CTA and the CTA-pair dataflow form (LBIC_FLOW_PAIR_MIN_ROWS), default bench at several batch sizes
mkdir -p gpurun_out
L=gpurun_out/r2_midsize2.log
: > $L
run() {
  echo "== pair_min_rows=$1 images=$2" >> $L
  LBIC_FLOW_PAIR_MIN_ROWS=$1 timeout 600 python bench.py --images $2 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-reference-container 2>> $L | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.1f enc %.1f dec %.1f sm_mhz %s identical %s' % (d['value'], d['encode_mpix_s'], d['decode_mpix_s'], d['clocks']['sm_mhz'], d['enc_dec_identical']))" >> $L
}
for n in 96 128 192 256 1024; do
  for t in 4096 8192 12288; do
    run $t $n
  done
done
cat $L
