#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --images 512 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/prof4_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_ws -s 1500 -c 3 -o gpurun_out/prof_ws_r1 $CMD > gpurun_out/ncu_ws.log 2>&1
tail -2 gpurun_out/ncu_ws.log
for cfg in B4_highrate B8_highrate B16_lowrate; do
  echo "== $cfg" >> gpurun_out/prof4_cfg.log
  timeout 600 python bench.py --config $cfg --images 64 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline >> gpurun_out/prof4_cfg.log 2>&1
done
