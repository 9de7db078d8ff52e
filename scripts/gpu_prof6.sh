#!/bin/bash
mkdir -p gpurun_out
python scripts/decode_launches.py 1024 0 > gpurun_out/prof6_plain.log 2>&1 || { tail -5 gpurun_out/prof6_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 6000 --csv --log-file gpurun_out/launches_r1_decode1024.csv python scripts/decode_launches.py 1024 0 > gpurun_out/ncu_launch6.log 2>&1
tail -2 gpurun_out/ncu_launch6.log
