#!/bin/bash
# chunked tile order of the dataflow launch (LBIC_FLOW_CHUNK = row blocks of 256 rows per chunk; 0 = one chunk): default bench
mkdir -p gpurun_out
L=gpurun_out/r2_chunk.log
: > $L
for c in 0 96 64 48 40 32 24 0; do
  echo "== LBIC_FLOW_CHUNK=$c" >> $L
  LBIC_FLOW_CHUNK=$c timeout 600 python bench.py --images 1024 --steps 2 --warmup 2 --no-e2e --no-cpu-baseline --no-reference-container 2>> $L | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.1f enc %.1f dec %.1f sm_mhz %s identical %s' % (d['value'], d['encode_mpix_s'], d['decode_mpix_s'], d['clocks']['sm_mhz'], d['enc_dec_identical']))" >> $L
done
cat $L
