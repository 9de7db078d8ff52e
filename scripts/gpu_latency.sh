#!/bin/bash
mkdir -p gpurun_out
python - <<'PY' > gpurun_out/latency.log 2>&1
import sys, os, time, torch
sys.path.insert(0, os.getcwd())
import lbic_b200
from lbic_b200 import weights
from lbic_b200.layout import arrange_block_pixels_to_channel_dim
from lbic_b200.net import BlockBasedImgCompLossyNetv9
dev = torch.device("cuda:0")
def run(cfgname, n, H, W, lanes_list=(0, 1)):
    cfg = lbic_b200.load_config(cfgname)
    m = BlockBasedImgCompLossyNetv9(cfg, device=dev)
    m.load_state_dict(weights.synth_state_dict(cfg, 1337)); m.update(force=True)
    B = cfg.block_size
    img = torch.rand(n, 3, H, W, device=dev)
    x = arrange_block_pixels_to_channel_dim(img - 0.5, B); del img
    for lanes in lanes_list:
        out = m.encode_device(x, lanes=lanes); torch.cuda.synchronize()
        t = time.perf_counter(); out = m.encode_device(x, lanes=lanes, out=out); torch.cuda.synchronize(); te = time.perf_counter() - t
        z = m.decode_device(out.streams, out.lens, n, H // B, W // B, lanes=lanes); torch.cuda.synchronize()
        t = time.perf_counter(); z = m.decode_device(out.streams, out.lens, n, H // B, W // B, lanes=lanes); torch.cuda.synchronize(); td = time.perf_counter() - t
        print(f"{cfgname:12s} n={n:4d} {W}x{H} lanes={lanes}: encode {te*1e3:8.1f} ms ({n*H*W/te/1e6:7.1f} Mpix/s)  decode {td*1e3:8.1f} ms ({n*H*W/td/1e6:7.1f} Mpix/s)  identical={bool(torch.equal(z, out.zhat))}", flush=True)
    del m, x, out, z
    torch.cuda.empty_cache()
for n in (1, 8, 24, 64):
    run("B8_lowrate", n, 512, 768)
run("B4_highrate", 24, 512, 768)
run("B8_highrate", 24, 512, 768)
run("B8_highrate", 256, 512, 768, (0,))
run("B16_lowrate", 8, 2048, 2048, (0,))
run("B8_lowrate", 1, 8192, 8192, (0,))
PY
cat gpurun_out/latency.log
