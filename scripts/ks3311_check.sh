#!/bin/bash
for c in B4_highrate B8_highrate; do
    timeout 300 python bench.py --config $c --images 256 --steps 1 --warmup 1 --no-cpu-baseline 2>/dev/null | tail -1 > /tmp/line.json
    python - "$c" <<'PY'
import sys, json
d = json.loads(open('/tmp/line.json').read())
print(sys.argv[1], "enc", round(d["encode_mpix_s"]), "dec", round(d["decode_mpix_s"]), "identical", d["enc_dec_identical"], "e2e parity", d["e2e"]["parity"], "ref container identical", d["reference_container"]["enc_dec_identical"])
PY
done
