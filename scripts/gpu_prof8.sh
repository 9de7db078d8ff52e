#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 6000 --csv --log-file gpurun_out/launches_r1_encode1024.csv python scripts/decode_launches.py 1024 0 enc > gpurun_out/ncu_launch8.log 2>&1
tail -2 gpurun_out/ncu_launch8.log
