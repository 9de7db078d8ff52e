#!/bin/bash
LBIC_FLOW_CHUNK=2 timeout 300 python -m pytest tests -m gpu -x -q -k "dataflow" 2>&1 | tail -2
for ch in 0 16 32 64; do
    LBIC_FLOW_CHUNK=$ch timeout 300 python bench.py --images ${N:-1024} --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-reference-container 2>/dev/null | tail -1 > /tmp/line.json
    python - "$ch" <<'PY'
import sys, json
d = json.loads(open('/tmp/line.json').read())
print("chunk", sys.argv[1], d["config"].get("images_per_gpu"), round(d["encode_mpix_s"]), round(d["decode_mpix_s"]), round(d["value"]), d["clocks"]["sm_mhz"])
PY
done
