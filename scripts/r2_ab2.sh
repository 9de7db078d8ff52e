#!/bin/bash
# equivalence tests with the current build, then A/B of scripts/_ab/{old,new}.so (alternating) on the default bench
mkdir -p gpurun_out
L=gpurun_out/r2_ab2.log
SO=learned-block-based-image-compression_b200/liblbic_b200.so
cp $SO /tmp/cur.so
echo "== tests (current build)" > $L
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -k "gemm_core or wave_kernel or warp_specialised or dataflow or encode_matches or decode_roundtrip or full_size_fixed" >> $L 2>&1
echo "exit $?" >> $L
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
cat $L
