#!/bin/bash
# default bench on all 8 GPUs of one box (weak scaling, 1024 images per GPU), launched as the driver does
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 3 --warmup 3 --refc-images 0 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.log
echo "exit $?"
cut -c1-300 gpurun_out/r2_bench_n8.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n8.json').read().strip().splitlines()[-1])
print('value %.1f e2e %.1f ratio %.3f enc %.1f dec %.1f refc %s clocks %s' % (d['value'], d['e2e']['value'], d['e2e']['ratio_to_device_value'], d['encode_mpix_s'], d['decode_mpix_s'], d['reference_container']['value'] if d.get('reference_container') else None, d['clocks']))
PY
