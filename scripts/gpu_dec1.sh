#!/bin/bash
mkdir -p gpurun_out
python - <<'PY' > gpurun_out/dec1.log 2>&1
import sys, os, time, torch
sys.path.insert(0, os.getcwd())
import lbic_b200
from lbic_b200 import weights
from lbic_b200.layout import arrange_block_pixels_to_channel_dim
from lbic_b200.net import BlockBasedImgCompLossyNetv9
dev = torch.device("cuda:0")
cfg = lbic_b200.load_config("B8_lowrate")
m = BlockBasedImgCompLossyNetv9(cfg, device=dev)
m.load_state_dict(weights.synth_state_dict(cfg, 1337)); m.update(force=True)
for n in (256, 1024):
    img = torch.rand(n, 3, 512, 768, device=dev)
    x = arrange_block_pixels_to_channel_dim(img - 0.5, 8); del img
    out = m.encode_device(x, lanes=1); torch.cuda.synchronize()
    t = time.perf_counter(); out = m.encode_device(x, lanes=1, out=out); torch.cuda.synchronize(); te = time.perf_counter() - t
    print(f"n={n} lanes=1 encode {te*1e3:.0f} ms ({n*512*768/te/1e6:.1f} Mpix/s)", flush=True)
    for chain, S in ((0, 0), (1, 0), (1, 4), (1, 8)):
        m.set_option("chain", chain); m.set_option("cluster", S)
        t = time.perf_counter(); z = m.decode_device(out.streams, out.lens, n, 64, 96, lanes=1); torch.cuda.synchronize(); td = time.perf_counter() - t
        print(f"n={n} lanes=1 chain={chain} S={S}: decode {td*1e3:.0f} ms ({n*512*768/td/1e6:.1f} Mpix/s) identical={bool(torch.equal(z, out.zhat))}", flush=True)
    m.set_option("chain", 0); m.set_option("cluster", 0)
    del x, out, z
PY
cat gpurun_out/dec1.log
