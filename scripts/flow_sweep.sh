#!/bin/bash
for n in 64 128 256 512; do
  for f in 0 1; do
    LBIC_FLOW=$f LBIC_FLOW_MIN_ROWS=${MINROWS:-1024} timeout 300 python bench.py --images $n --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-reference-container 2>/dev/null | tail -1 > /tmp/line.json
    python - "$f" <<'PY'
import sys, json
d = json.loads(open('/tmp/line.json').read())
print("flow", sys.argv[1], d["config"].get("images_per_gpu"), round(d["encode_mpix_s"]), round(d["decode_mpix_s"]), round(d["value"]), d["clocks"]["sm_mhz"], d["gpu_launches"])
PY
  done
done
