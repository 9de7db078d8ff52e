#!/bin/bash
# warp-uniform MMA issue (elect_one): kernel-equivalence tests, one-image latency, default bench
mkdir -p gpurun_out
L=gpurun_out/r2_check7.log
echo "== equivalence tests" > $L
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -k "gemm_core or wave_kernel or warp_specialised or dataflow or encode_matches or decode_roundtrip" >> $L 2>&1
echo "exit $?" >> $L
echo "== latency" >> $L
timeout 600 python scripts/latency.py --quick > gpurun_out/r2_latency_v6.jsonl 2>> $L
echo "exit $?" >> $L
echo "== bench" >> $L
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r2_bench_v6.json 2>> $L
echo "exit $?" >> $L
grep -E "^exit|passed|failed|^==|Error|^FAILED" $L
cat gpurun_out/r2_latency_v6.jsonl; cut -c1-1500 gpurun_out/r2_bench_v6.json
