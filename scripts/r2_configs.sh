#!/bin/bash
# BASELINE.json configs 2-4 through bench.py on N GPUs (usage: r2_configs.sh N).  Lines are kept under gpurun_out/ and
# copied to profiles/r2_bench_configs.jsonl.
N=$1
mkdir -p gpurun_out
OUT=gpurun_out/r2_bench_configs_n$N.jsonl
L=gpurun_out/r2_bench_configs_n$N.log
: > $OUT; : > $L
run() {
  if [ "$N" = "1" ]; then
    timeout 900 python bench.py --gpus 1 --steps 2 --warmup 3 --no-cpu-baseline "$@" >> $OUT 2>> $L
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 2 --warmup 3 --no-cpu-baseline "$@" >> $OUT 2>> $L
  fi
  echo "exit $? $*" >> $L
}
# C3: B8_highrate, 1024 images of 768x512 image-sharded over N GPUs (strong scaling)
run --config B8_highrate --scaling strong --total-images 1024 --no-reference-container
if [ "$N" = "1" ]; then
  # C2: B4_highrate, batch of 24 images of 768x512 on one GPU
  run --config B4_highrate --images 24
  # C4 per-GPU share: B16_lowrate, 8 images of 2048x2048
  run --config B16_lowrate --height 2048 --width 2048 --images 8 --no-reference-container
fi
if [ "$N" = "8" ]; then
  # C4: B16_lowrate 2048x2048 on 8 GPUs, 8 images per GPU (64 in total)
  run --config B16_lowrate --height 2048 --width 2048 --images 8 --no-reference-container
fi
grep -E "^exit" $L
python - <<'PY'
import json,sys,glob
for f in sorted(glob.glob("gpurun_out/r2_bench_configs_n*.jsonl")):
    for l in open(f):
        if l.startswith("{"):
            d=json.loads(l); print(f, d["n_gpus"], d["scaling"], d["config"]["workload"][:60], "value %.1f e2e %.1f enc %.1f dec %.1f"%(d["value"], d["e2e"]["value"] if d.get("e2e") else -1, d["encode_mpix_s"], d["decode_mpix_s"]))
PY
