#!/bin/bash
# A/B of two builds of liblbic_b200.so on one box (scripts/_ab/old.so, new.so), alternating; then the wave kernel's ring depth
mkdir -p gpurun_out
L=gpurun_out/r2_ab.log
SO=learned-block-based-image-compression_b200/liblbic_b200.so
cp $SO /tmp/cur.so
: > $L
for round in 1 2; do
  for v in old new; do
    cp scripts/_ab/$v.so $SO
    for n in 1024 256; do
      echo "== $v images=$n round=$round" >> $L
      timeout 600 python bench.py --images $n --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-reference-container 2>> $L | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.1f enc %.1f dec %.1f sm_mhz %s' % (d['value'], d['encode_mpix_s'], d['decode_mpix_s'], d['clocks']['sm_mhz']))" >> $L
    done
  done
done
cp /tmp/cur.so $SO
for st in 2 3 4 12; do
  echo "== wave ring depth cap $st" >> $L
  LBIC_WAVE_STAGES=$st timeout 300 python scripts/latency.py --quick 2>> $L | grep '"lane", "wave_kernel": true' | cut -c1-200 >> $L
done
cat $L
