#!/bin/bash
# new dataflow thresholds (single-CTA form 2560-8191 rows, CTA pairs from 8192): GPU test suite, then mid-size benches
mkdir -p gpurun_out
L=gpurun_out/r2_check8.log
echo "== pytest gpu" > $L
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 >> $L 2>&1
echo "exit $?" >> $L
run() {
  echo "== $*" >> $L
  timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-reference-container "$@" 2>> $L | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.1f enc %.1f dec %.1f sm_mhz %s identical %s' % (d['value'], d['encode_mpix_s'], d['decode_mpix_s'], d['clocks']['sm_mhz'], d['enc_dec_identical']))" >> $L
}
run --images 128
run --images 256
run --config B8_highrate --images 128
LBIC_FLOW_MIN_ROWS=4096 LBIC_FLOW_PAIR_MIN_ROWS=4096 run --config B8_highrate --images 128
run --config B16_lowrate --height 2048 --width 2048 --images 8
LBIC_FLOW_MIN_ROWS=4096 LBIC_FLOW_PAIR_MIN_ROWS=4096 run --config B16_lowrate --height 2048 --width 2048 --images 8
grep -E "^exit|^==|passed|failed|value" $L
