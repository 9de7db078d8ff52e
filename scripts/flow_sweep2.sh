#!/bin/bash
for mx in 16384 24576 32768 40960 1000000; do
    LBIC_FLOW=1 LBIC_FLOW_MIN_ROWS=8192 LBIC_FLOW_MAX_ROWS=$mx timeout 300 python bench.py --images 1024 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-reference-container 2>/dev/null | tail -1 > /tmp/line.json
    python - "$mx" <<'PY'
import sys, json
d = json.loads(open('/tmp/line.json').read())
print("max_rows", sys.argv[1], d["config"].get("images_per_gpu"), round(d["encode_mpix_s"]), round(d["decode_mpix_s"]), round(d["value"]), d["clocks"]["sm_mhz"], d["gpu_launches"])
PY
done
