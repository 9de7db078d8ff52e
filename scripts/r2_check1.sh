#!/bin/bash
# round 2, GPU call 1: full GPU test suite, GEMM precision experiment, default bench, full-size parity sweep
mkdir -p gpurun_out
L=gpurun_out/r2_check1.log
echo "== pytest gpu" > $L
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 1200 >> $L 2>&1
echo "exit $?" >> $L
echo "== smoke" >> $L
timeout 300 python __graft_entry__.py smoke >> $L 2>&1
echo "exit $?" >> $L
echo "== gemm precision" >> $L
timeout 300 python scripts/gemm_precision.py > gpurun_out/r2_gemm_precision.jsonl 2>> $L
echo "exit $?" >> $L
echo "== bench default" >> $L
( time timeout 1200 python bench.py ) > gpurun_out/r2_bench_default.json 2>> $L
echo "exit $?" >> $L
echo "== parity sweep" >> $L
timeout 1200 python scripts/parity_sweep.py > gpurun_out/r2_parity_sweep.jsonl 2>> $L
echo "exit $?" >> $L
grep -E "^exit|passed|failed|^==|Error|^real|smoke:" $L
