#!/bin/bash
for n in 128 192 256 384; do
  for mr in 4096 6144 8192 1000000; do
    LBIC_FLOW_MIN_ROWS=$mr timeout 300 python bench.py --images $n --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-reference-container 2>/dev/null | tail -1 > /tmp/line.json
    python - "$mr" <<'PY'
import sys, json
d = json.loads(open('/tmp/line.json').read())
print("min_rows", sys.argv[1], "images", d["config"]["images_per_gpu"], "enc", round(d["encode_mpix_s"]), "dec", round(d["decode_mpix_s"]), "rt", round(d["value"]), "MHz", d["clocks"]["sm_mhz"])
PY
  done
done
