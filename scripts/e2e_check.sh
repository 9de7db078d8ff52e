#!/bin/bash
timeout 600 python -m pytest tests -m gpu -x -q -k "golden or roundtrip or error or capacity or image_codec or batch_invariance" 2>&1 | tail -2
for n in 256 1024; do
timeout 400 python bench.py --images $n --steps 2 --warmup 1 --no-cpu-baseline --no-reference-container 2>/dev/null | tail -1 > /tmp/line.json
python - <<'PY'
import json
d = json.loads(open('/tmp/line.json').read())
print("images", d["config"]["images_per_gpu"], "device rt", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["e2e"]["parity"], "enc", round(d["encode_mpix_s"]), "dec", round(d["decode_mpix_s"]))
PY
done
