#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
echo "== reference arm" > gpurun_out/multi_$N.log
timeout 600 python bench.py --impl reference --gpus 1 --steps 1 --warmup 0 --ref-blocks 768 >> gpurun_out/multi_$N.log 2>&1
echo "== bench N=1" >> gpurun_out/multi_$N.log
timeout 900 python bench.py --gpus 1 --steps 3 --warmup 3 >> gpurun_out/multi_$N.log 2>&1
echo "exit $?" >> gpurun_out/multi_$N.log
echo "== bench N=$N" >> gpurun_out/multi_$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 >> gpurun_out/multi_$N.log 2>&1
echo "exit $?" >> gpurun_out/multi_$N.log
grep -E "^exit|^==|Error" gpurun_out/multi_$N.log
