#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_ab3.log
SO=learned-block-based-image-compression_b200/liblbic_b200.so
cp $SO /tmp/cur.so
: > $L
cat > /tmp/rc.py <<'PY'
import os, sys, json, torch
sys.path.insert(0, os.getcwd())
import lbic_b200
from lbic_b200 import weights
from lbic_b200.layout import arrange_block_pixels_to_channel_dim
from lbic_b200.net import BlockBasedImgCompLossyNetv9
dev = torch.device("cuda:0")
cfg = lbic_b200.load_config("B8_lowrate")
m = BlockBasedImgCompLossyNetv9(cfg, device=dev)
m.load_state_dict(weights.synth_state_dict(cfg, 1337)); m.update(force=True)
n = 1024
g = torch.Generator(device=dev); g.manual_seed(5)
x = arrange_block_pixels_to_channel_dim(torch.rand(n, 3, 512, 768, device=dev, generator=g) - 0.5, 8)
enc = m.encode_device(x, lanes=1)
z = m.decode_device(enc.streams, enc.lens, n, 64, 96, lanes=1)
best = 1e9
for _ in range(2):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    z = m.decode_device(enc.streams, enc.lens, n, 64, 96, lanes=1)
    b.record(); torch.cuda.synchronize()
    best = min(best, a.elapsed_time(b))
print("reference container decode, 1024 images: %.1f ms = %.1f Mpixel/s" % (best, n * 512 * 768 / best / 1e3))
x1 = x[:1].contiguous()
e1 = m.encode_device(x1, lanes=0)
best = 1e9
for _ in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    e1 = m.encode_device(x1, lanes=0, out=e1)
    b.record(); torch.cuda.synchronize()
    best = min(best, a.elapsed_time(b))
print("one image encode (wave kernel): %.2f ms" % best)
PY
for round in 1 2; do
  for v in old new; do
    cp scripts/_ab/$v.so $SO
    echo "== $v round=$round" >> $L
    timeout 300 python /tmp/rc.py >> $L 2>&1
  done
done
cp /tmp/cur.so $SO
cat $L
