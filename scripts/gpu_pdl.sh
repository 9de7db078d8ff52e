#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pdl_tests.log 2>&1; echo "exit $?" >> gpurun_out/pdl_tests.log; tail -3 gpurun_out/pdl_tests.log
python - <<'PY' > gpurun_out/pdl.log 2>&1
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
for n in (64, 256, 1024):
    img = torch.rand(n, 3, 512, 768, device=dev)
    x = arrange_block_pixels_to_channel_dim(img - 0.5, 8); del img
    out = m.encode_device(x, lanes=0)
    for pdl in (0, 1):
        m.set_option("pdl", pdl)
        m.encode_device(x, lanes=0, out=out); torch.cuda.synchronize()
        t = time.perf_counter(); m.encode_device(x, lanes=0, out=out); torch.cuda.synchronize(); te = time.perf_counter() - t
        t = time.perf_counter(); z = m.decode_device(out.streams, out.lens, n, 64, 96, lanes=0); torch.cuda.synchronize(); td = time.perf_counter() - t
        print(f"n={n} pdl={pdl}: encode {te*1e3:.1f} ms ({n*512*768/te/1e6:.1f} Mpix/s) decode {td*1e3:.1f} ms ({n*512*768/td/1e6:.1f} Mpix/s) identical={bool(torch.equal(z, out.zhat))}", flush=True)
    if n == 256:
        o1 = m.encode_device(x, lanes=1)
        for pdl in (0, 1):
            m.set_option("pdl", pdl)
            t = time.perf_counter(); z = m.decode_device(o1.streams, o1.lens, n, 64, 96, lanes=1); torch.cuda.synchronize(); td = time.perf_counter() - t
            print(f"n={n} lanes=1 pdl={pdl}: decode {td*1e3:.0f} ms ({n*512*768/td/1e6:.1f} Mpix/s) identical={bool(torch.equal(z, o1.zhat))}", flush=True)
        del o1
    del x, out, z
PY
cat gpurun_out/pdl.log
