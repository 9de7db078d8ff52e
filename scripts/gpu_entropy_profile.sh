#!/bin/bash
# ncu --set full of the two entropy-stage kernels inside a 1024-image encode / decode
mkdir -p gpurun_out
ncu --set full --clock-control none --profile-from-start off -k regex:rans_encode_thread -c 1 -o gpurun_out/r1_rans_enc -f python scripts/decode_launches.py 1024 0 enc > gpurun_out/ncu_rans_enc.log 2>&1
ncu --set full --clock-control none --profile-from-start off -k regex:rans_dec_step_thread -s 110 -c 1 -o gpurun_out/r1_rans_dec -f python scripts/decode_launches.py 1024 0 > gpurun_out/ncu_rans_dec.log 2>&1
tail -1 gpurun_out/ncu_rans_enc.log; tail -1 gpurun_out/ncu_rans_dec.log
