#!/bin/bash
# what bounds E1 of a KS3311 step (K = 5 E1 = 5760, 33 MB of weights per step): ring depth and tile width probes
mkdir -p gpurun_out
L=gpurun_out/r2_e1_probe.log
: > $L
run() {
  echo "== $*" >> $L
  env "$@" LBIC_TRACE_CONFIG=B8_highrate bash scripts/r2_trace.sh > /dev/null 2>&1
  grep -E "^step|^  E[0-3]|^  G0|^  GATHER5" gpurun_out/wave_trace_enc_summary.txt >> $L
}
run LBIC_WAVE_STAGES=6
run LBIC_WAVE_STAGES=3
run LBIC_WAVE_STAGES=2
run LBIC_WAVE_BN=64
run LBIC_WAVE_XG=48
cat $L
