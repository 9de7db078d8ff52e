#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for n in 1 8 24 64; do
  for f in 0 1; do
    LBIC_FLOW_SMALL=$f timeout 300 python bench.py --images $n --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-reference-container 2>/dev/null | tail -1 > /tmp/line.json
    python - "$f" <<'PY'
import sys, json
d = json.loads(open('/tmp/line.json').read())
n = d["config"]["images_per_gpu"]; px = n * 512 * 768 / 1e6
print("flow_small", sys.argv[1], "images", n, "enc ms", round(px / d["encode_mpix_s"] * 1e3, 2), "dec ms", round(px / d["decode_mpix_s"] * 1e3, 2), "enc", round(d["encode_mpix_s"], 1), "dec", round(d["decode_mpix_s"], 1), "launches", d["gpu_launches"])
PY
  done
done
