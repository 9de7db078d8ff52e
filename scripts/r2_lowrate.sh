#!/bin/bash
# the default bench with a synthetic model of realistic rate (~1 bpp instead of ~11): what the entropy stage and the
# stream copies cost then
mkdir -p gpurun_out
L=gpurun_out/r2_lowrate.log
: > $L
for g in "60 5.2" "6 3.0" "2.5 2.0"; do
  set -- $g
  echo "== latent_gain $1 scale_span $2" >> $L
  timeout 900 python bench.py --latent-gain $1 --scale-span $2 --no-cpu-baseline --refc-images 0 2>> $L | python -c "
import json,sys
d=json.loads(sys.stdin.read()); rc=d['reference_container']
print('bpp %.2f value %.1f enc %.1f dec %.1f e2e %.1f (%.3f) refc %.1f (enc %.1f dec %.1f) identical %s e2e parity %s' % (d['bpp'], d['value'], d['encode_mpix_s'], d['decode_mpix_s'], d['e2e']['value'], d['e2e']['ratio_to_device_value'], rc['value'], rc['encode_mpix_s'], rc['decode_mpix_s'], d['enc_dec_identical'], d['e2e']['parity']))" >> $L
done
cat $L
