#!/bin/bash
# QUANT epilogue with staged side inputs: tests, isolated timing, one-image latency, bench A/B
mkdir -p gpurun_out
L=gpurun_out/r2_quant.log
SO=learned-block-based-image-compression_b200/liblbic_b200.so
cp $SO /tmp/cur.so
echo "== tests" > $L
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -k "encode_matches or decode_roundtrip or wave_kernel or warp_specialised or dataflow or batch_invariance or full_size_fixed" >> $L 2>&1
rc=$?
echo "exit $rc" >> $L
if [ $rc -eq 0 ]; then
for v in old new; do
  cp scripts/_ab/$v.so $SO
  echo "== $v: isolated QUANT launches" >> $L
  timeout 300 python scripts/epi_modes.py 24576 2>&1 | grep "pair=1" | grep -E "QUANT" >> $L
  echo "== $v: one image" >> $L
  timeout 300 python scripts/latency.py --quick 2>> $L | grep '"lane", "wave_kernel": true' | cut -c1-190 >> $L
done
for round in 1 2; do
  for v in old new; do
    cp scripts/_ab/$v.so $SO
    echo "== $v images=1024 round=$round" >> $L
    timeout 600 python bench.py --images 1024 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-reference-container 2>> $L | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.1f enc %.1f dec %.1f sm_mhz %s F3 %s' % (d['value'], d['encode_mpix_s'], d['decode_mpix_s'], d['clocks']['sm_mhz'], d['roofline']['layers_tflops_per_layer_launches']['F3']))" >> $L
  done
done
fi
cp /tmp/cur.so $SO
cat $L
