#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_check5.log
echo "== new tests" > $L
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -k "postprocessing or metrics or band or eval_model or codec" >> $L 2>&1
echo "exit $?" >> $L
echo "== pytest gpu (all)" >> $L
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 >> $L 2>&1
echo "exit $?" >> $L
grep -E "^exit|passed|failed|^==|Error|^FAILED" $L
