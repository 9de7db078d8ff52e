#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/tr.py <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
import lbic_b200
from lbic_b200 import weights
from lbic_b200.layout import arrange_block_pixels_to_channel_dim
from lbic_b200.net import BlockBasedImgCompLossyNetv9
dev = torch.device("cuda:0")
cfg = lbic_b200.load_config(os.environ.get("LBIC_TRACE_CONFIG", "B8_lowrate"))
m = BlockBasedImgCompLossyNetv9(cfg, device=dev)
m.load_state_dict(weights.synth_state_dict(cfg, 1337)); m.update(force=True)
x = arrange_block_pixels_to_channel_dim(torch.rand(1, 3, 512, 768, device=dev) - 0.5, 8)
mode = sys.argv[1]
if mode == "dec":
    m.set_option("wave", 0)
o = m.encode_device(x, lanes=0)
torch.cuda.synchronize()
if mode == "dec":
    m.set_option("wave", 1)
    m.decode_device(o.streams, o.lens, 1, 64, 96, lanes=0)
torch.cuda.synchronize()
print("done", mode)
PY
# the trace is taken on the FIRST launch that covers the step: encode when tracing "enc", else skip encode's by env order
LBIC_WAVE_TRACE=110 LBIC_WAVE_TRACE_STEPS=2 LBIC_WAVE_TRACE_FILE=gpurun_out/wave_trace_enc.txt python /tmp/tr.py enc > gpurun_out/r2_trace.log 2>&1
python scripts/wave_trace.py gpurun_out/wave_trace_enc.txt > gpurun_out/wave_trace_enc_summary.txt 2>&1
LBIC_WAVE_TRACE=110 LBIC_WAVE_TRACE_STEPS=2 LBIC_WAVE_TRACE_FILE=gpurun_out/wave_trace_dec.txt python /tmp/tr.py dec >> gpurun_out/r2_trace.log 2>&1
python scripts/wave_trace.py gpurun_out/wave_trace_dec.txt > gpurun_out/wave_trace_dec_summary.txt 2>&1
cat gpurun_out/wave_trace_enc_summary.txt
