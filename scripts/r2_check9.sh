#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_check9.log
: > $L
run() {
  echo "== $*" >> $L
  timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-reference-container "$@" 2>> $L | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.1f enc %.1f dec %.1f sm_mhz %s identical %s' % (d['value'], d['encode_mpix_s'], d['decode_mpix_s'], d['clocks']['sm_mhz'], d['enc_dec_identical']))" >> $L
}
for n in 64 128 256; do
  echo "-- new thresholds" >> $L
  run --config B8_highrate --images $n
  echo "-- old thresholds (4096 / pairs only)" >> $L
  LBIC_FLOW_MIN_ROWS=4096 LBIC_FLOW_PAIR_MIN_ROWS=1 run --config B8_highrate --images $n
  echo "-- new lower threshold, pairs only" >> $L
  LBIC_FLOW_PAIR_MIN_ROWS=1 run --config B8_highrate --images $n
done
cat $L
