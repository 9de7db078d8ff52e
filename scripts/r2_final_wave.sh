#!/bin/bash
# final state of the wave kernel: latencies (one image, small batches, large images; all topologies, both containers) and
# ncu --set full of the KS3311 encode launch
mkdir -p gpurun_out
timeout 600 python scripts/latency.py > gpurun_out/r2_latency_wave_v8.jsonl 2> gpurun_out/r2_final_wave.log
echo "latency exit $?" >> gpurun_out/r2_final_wave.log
timeout 300 python scripts/latency_topologies.py > gpurun_out/r2_latency_topologies_v8.jsonl 2>> gpurun_out/r2_final_wave.log
echo "topologies exit $?" >> gpurun_out/r2_final_wave.log
LBIC_TRACE_CONFIG=B8_highrate timeout 120 python scripts/one_image_wave.py enc >> gpurun_out/r2_final_wave.log 2>&1 && \
LBIC_TRACE_CONFIG=B8_highrate timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_wave -c 1 -f -o gpurun_out/r2_wave_k3 python scripts/one_image_wave.py enc > gpurun_out/r2_ncu_wave_k3.log 2>&1
echo "ncu exit $?" >> gpurun_out/r2_final_wave.log
cat gpurun_out/r2_final_wave.log gpurun_out/r2_latency_wave_v8.jsonl gpurun_out/r2_latency_topologies_v8.jsonl | cut -c1-330
