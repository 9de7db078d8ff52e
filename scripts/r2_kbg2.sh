#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_kbg2.log
echo "== tests" > $L
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -k "wave_kernel or encode_matches or dataflow or warp_specialised" >> $L 2>&1
echo "exit $?" >> $L
timeout 300 python scripts/latency.py --quick 2>> $L | grep '"wave_kernel": true' | cut -c1-200 >> $L
bash scripts/r2_trace.sh > /dev/null 2>&1
tail -20 gpurun_out/wave_trace_enc_summary.txt >> $L
cat $L
