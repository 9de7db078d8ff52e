#!/bin/bash
# same-box A/B (scripts/_ab/old.so against new.so) on all topologies, then ncu --set full of the KS3311 encode launch with new.so
LBIC_LAT_CONFIGS=B8_lowrate,B8_highrate,B4_highrate,B16_lowrate bash scripts/r2_wave_ab_quick.sh > /dev/null 2>&1
cp gpurun_out/r2_wave_ab_quick.log gpurun_out/r2_wave_gather_ab.log
LBIC_TRACE_CONFIG=B8_highrate timeout 120 python scripts/one_image_wave.py enc >> gpurun_out/r2_wave_gather_ab.log 2>&1 && \
LBIC_TRACE_CONFIG=B8_highrate timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_wave -c 1 -f -o gpurun_out/r2_wave_k3 python scripts/one_image_wave.py enc > gpurun_out/r2_ncu_wave_k3.log 2>&1
echo "ncu exit $?" >> gpurun_out/r2_wave_gather_ab.log
cat gpurun_out/r2_wave_gather_ab.log
