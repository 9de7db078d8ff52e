#!/bin/bash
# N-GPU band pipeline: correctness at 2048^2, numbers at 8192^2 (BASELINE config 5).  usage: r2_band_n.sh N
N=$1
mkdir -p gpurun_out
L=gpurun_out/r2_band_n$N.log
echo "== band 2048 on $N GPUs" > $L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/band_bench.py --size 2048 > gpurun_out/r2_band_n${N}_2048.json 2>> $L
echo "exit $?" >> $L
echo "== band 8192 on $N GPUs" >> $L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 scripts/band_bench.py --size 8192 --reps 1 > gpurun_out/r2_band_n${N}_8192.json 2>> $L
echo "exit $?" >> $L
grep -E "^exit|^==|Error" $L; grep -h "^{" gpurun_out/r2_band_n${N}_2048.json gpurun_out/r2_band_n${N}_8192.json
