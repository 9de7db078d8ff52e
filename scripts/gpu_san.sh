#!/bin/bash
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python __graft_entry__.py smoke > gpurun_out/memcheck_smoke.log 2>&1
echo "memcheck exit $?" >> gpurun_out/memcheck_smoke.log
tail -6 gpurun_out/memcheck_smoke.log
python - <<'PY' > gpurun_out/big_image.log 2>&1
import sys, os, time, torch
sys.path.insert(0, os.getcwd())
import lbic_b200
from lbic_b200 import weights
from lbic_b200.layout import arrange_block_pixels_to_channel_dim
from lbic_b200.net import BlockBasedImgCompLossyNetv9
dev = torch.device("cuda:0")
for cfgname, n, H, W in (("B16_lowrate", 8, 2048, 2048), ("B8_lowrate", 1, 4096, 4096)):
    cfg = lbic_b200.load_config(cfgname)
    m = BlockBasedImgCompLossyNetv9(cfg, device=dev)
    m.load_state_dict(weights.synth_state_dict(cfg, 1337)); m.update(force=True)
    B = cfg.block_size
    img = torch.rand(n, 3, H, W, device=dev)
    x = arrange_block_pixels_to_channel_dim(img - 0.5, B); del img
    for lanes in (0, 1):
        out = m.encode_device(x, lanes=lanes); torch.cuda.synchronize()
        t = time.perf_counter(); out = m.encode_device(x, lanes=lanes, out=out); torch.cuda.synchronize(); te = time.perf_counter() - t
        t = time.perf_counter(); z = m.decode_device(out.streams, out.lens, n, H // B, W // B, lanes=lanes); torch.cuda.synchronize(); td = time.perf_counter() - t
        print(f"{cfgname} {n}x{W}x{H} lanes={lanes}: encode {te*1e3:.0f} ms ({n*H*W/te/1e6:.1f} Mpix/s) decode {td*1e3:.0f} ms ({n*H*W/td/1e6:.1f} Mpix/s) identical={bool(torch.equal(z, out.zhat))} bytes={out.lens.sum().item()}", flush=True)
    del m, x, out, z
PY
cat gpurun_out/big_image.log
