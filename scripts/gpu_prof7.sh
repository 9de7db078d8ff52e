#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,launch__grid_size,smsp__inst_executed.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum --clock-control none --profile-from-start off -k regex:rans_dec -s 100 -c 30 --csv --log-file gpurun_out/dec_step_thread.csv python scripts/decode_launches.py 1024 0 > gpurun_out/ncu_launch7.log 2>&1
tail -2 gpurun_out/ncu_launch7.log
