#!/bin/bash
# GPU test suite, smoke() and a short default bench; log in gpurun_out/full_check.log
mkdir -p gpurun_out
L=gpurun_out/full_check.log
echo "== pytest gpu" > $L
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 >> $L 2>&1
echo "exit $?" >> $L
echo "== smoke" >> $L
timeout 300 python __graft_entry__.py smoke >> $L 2>&1
echo "exit $?" >> $L
echo "== bench" >> $L
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/full_check_bench.json 2>> $L
echo "exit $?" >> $L
grep -E "^exit|^==|passed|failed|smoke:" $L
cut -c1-200 gpurun_out/full_check_bench.json
