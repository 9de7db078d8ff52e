#!/bin/bash
# MMA chain micro-benchmark, the decode tests touched by the shared-memory warp decode step, one-image latency, full GPU suite
mkdir -p gpurun_out
L=gpurun_out/r2_check6.log
echo "== mma chain" > $L
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mma_chain_bench scripts/mma_chain_bench.cu >> $L 2>&1 && timeout 120 /tmp/mma_chain_bench > gpurun_out/r2_mma_chain.txt 2>&1
echo "exit $?" >> $L
echo "== decode tests" >> $L
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -k "thread_per_stream or decode_roundtrip or wave_kernel" >> $L 2>&1
echo "exit $?" >> $L
echo "== latency" >> $L
timeout 600 python scripts/latency.py --quick > gpurun_out/r2_latency_v5.jsonl 2>> $L
echo "exit $?" >> $L
echo "== pytest gpu (all)" >> $L
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 >> $L 2>&1
echo "exit $?" >> $L
grep -E "^exit|passed|failed|^==|Error|^FAILED" $L
cat gpurun_out/r2_mma_chain.txt gpurun_out/r2_latency_v5.jsonl
