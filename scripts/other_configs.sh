#!/bin/bash
run() {
  timeout 400 python bench.py --config $1 --images $2 --height $3 --width $4 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-reference-container 2>/dev/null | tail -1 > /tmp/line.json
  python - "$@" <<'PY'
import sys, json
try:
    d = json.loads(open('/tmp/line.json').read())
    print("config", sys.argv[1], "images", sys.argv[2], f"{sys.argv[4]}x{sys.argv[3]}", "enc", round(d["encode_mpix_s"]), "dec", round(d["decode_mpix_s"]), "rt", round(d["value"]), "MHz", d["clocks"]["sm_mhz"], "TF/s", round(d["roofline"]["achieved"]))
except Exception as e:
    print("config", sys.argv[1:], "FAILED", e)
PY
}
run B8_highrate 512 512 768
run B4_highrate 24 512 768
run B4_highrate 512 512 768
run B16_lowrate 512 512 768
run B16_lowrate 8 2048 2048
