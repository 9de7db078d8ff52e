#!/bin/bash
# reference-container round trip vs batch size (device-resident): the raster-serial decode is 6144 steps whose time does not
# depend on the number of images, so throughput grows with the batch
mkdir -p gpurun_out
L=gpurun_out/r2_refc_batch.log
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
def timed(fn):
    best = 1e9
    for _ in range(2):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record(); out = fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best, out
for n in [int(v) for v in sys.argv[1:]]:
    g = torch.Generator(device=dev); g.manual_seed(5)
    x = arrange_block_pixels_to_channel_dim(torch.rand(n, 3, 512, 768, device=dev, generator=g) - 0.5, 8)
    for lanes in (1, 0):
        enc = m.encode_device(x, lanes=lanes)
        te, enc = timed(lambda: m.encode_device(x, lanes=lanes, out=enc))
        z = m.decode_device(enc.streams, enc.lens, n, 64, 96, lanes=lanes)
        td, z = timed(lambda: m.decode_device(enc.streams, enc.lens, n, 64, 96, lanes=lanes))
        px = n * 512 * 768
        print(json.dumps(dict(images=n, container="reference" if lanes else "lane", encode_mpix_s=round(px / te / 1e3, 1), decode_mpix_s=round(px / td / 1e3, 1),
                              round_trip_mpix_s=round(px / (te + td) / 1e3, 1), enc_dec_identical=bool(torch.equal(z, enc.zhat)),
                              mem_gb=round(torch.cuda.max_memory_allocated() / 2**30, 1))), flush=True)
        del enc, z
    del x
    torch.cuda.empty_cache()
PY
timeout 1200 python /tmp/rc.py 1024 2048 4096 >> $L 2>&1
cat $L
