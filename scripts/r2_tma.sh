#!/bin/bash
# TMA-store epilogue of the dataflow launch: equivalence tests, then A/B (LBIC_TMA_STORE=0/1, alternating) on the default bench
mkdir -p gpurun_out
L=gpurun_out/r2_tma.log
echo "== tests" > $L
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -k "dataflow or full_size_fixed or warp_specialised" >> $L 2>&1
rc=$?
echo "exit $rc" >> $L
if [ $rc -eq 0 ]; then
for round in 1 2; do
  for q in 0 1; do
    for n in 1024 256; do
      echo "== tma_store=$q images=$n round=$round" >> $L
      LBIC_TMA_STORE=$q timeout 600 python bench.py --images $n --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-reference-container 2>> $L | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.1f enc %.1f dec %.1f sm_mhz %s identical %s' % (d['value'], d['encode_mpix_s'], d['decode_mpix_s'], d['clocks']['sm_mhz'], d['enc_dec_identical']))" >> $L
    done
  done
done
fi
tail -40 $L
