#!/bin/bash
# grouped raster decode of the reference container: tests, then the reference-container round trip at 1 .. 4 groups
mkdir -p gpurun_out
L=gpurun_out/r2_groups.log
echo "== tests" > $L
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -k "image_groups or decode_roundtrip or thread_per_stream or stream_capacity or two_devices" >> $L 2>&1
echo "exit $?" >> $L
cat > /tmp/rc.py <<'PY'
import os, sys, time, json, torch
sys.path.insert(0, os.getcwd())
import lbic_b200
from lbic_b200 import weights
from lbic_b200.layout import arrange_block_pixels_to_channel_dim
from lbic_b200.net import BlockBasedImgCompLossyNetv9
dev = torch.device("cuda:0")
cfg = lbic_b200.load_config("B8_lowrate")
m = BlockBasedImgCompLossyNetv9(cfg, device=dev)
m.load_state_dict(weights.synth_state_dict(cfg, 1337)); m.update(force=True)
n = int(sys.argv[1])
g = torch.Generator(device=dev); g.manual_seed(5)
x = arrange_block_pixels_to_channel_dim(torch.rand(n, 3, 512, 768, device=dev, generator=g) - 0.5, 8)
enc = m.encode_device(x, lanes=1)
ref = None
for G in (1, 2, 3, 4):
    m.set_option("raster_groups", G)
    z = m.decode_device(enc.streams, enc.lens, n, 64, 96, lanes=1)
    best = 1e9
    for _ in range(2):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record()
        z = m.decode_device(enc.streams, enc.lens, n, 64, 96, lanes=1)
        b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    same = bool(torch.equal(z, enc.zhat))
    print(json.dumps(dict(images=n, groups=G, decode_ms=round(best, 1), decode_mpix_s=round(n * 512 * 768 / best / 1e3, 1), identical_to_encoder=same)), flush=True)
PY
for n in 1024 256; do
  echo "== reference container decode, $n images" >> $L
  timeout 600 python /tmp/rc.py $n >> $L 2>&1
done
cat $L
