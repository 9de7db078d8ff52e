#!/bin/bash
mkdir -p gpurun_out
echo "== ws test" > gpurun_out/check9.log
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "warp_specialised" >> gpurun_out/check9.log 2>&1
echo "exit $?" >> gpurun_out/check9.log
echo "== pytest gpu" >> gpurun_out/check9.log
timeout 1500 python -m pytest tests -m gpu -q >> gpurun_out/check9.log 2>&1
echo "exit $?" >> gpurun_out/check9.log
python - <<'PY' >> gpurun_out/check9.log 2>&1
import sys, os, time, torch
sys.path.insert(0, os.getcwd())
import lbic_b200
from lbic_b200 import weights
from lbic_b200.layout import arrange_block_pixels_to_channel_dim
from lbic_b200.net import BlockBasedImgCompLossyNetv9
cfg = lbic_b200.load_config("B8_lowrate")
dev = torch.device("cuda:0")
m = BlockBasedImgCompLossyNetv9(cfg, device=dev)
m.load_state_dict(weights.synth_state_dict(cfg, 1337)); m.update(force=True)
for n in (256, 512, 1024):
    img = torch.rand(n, 3, 512, 768, device=dev)
    x = arrange_block_pixels_to_channel_dim(img - 0.5, 8)
    del img
    out = m.encode_device(x, lanes=0)
    for ws in (0, 1, 2):
        m.set_option("ws", ws)
        m.encode_device(x, lanes=0, out=out); torch.cuda.synchronize()
        t = time.perf_counter()
        m.encode_device(x, lanes=0, out=out); torch.cuda.synchronize()
        dt = time.perf_counter() - t
        t = time.perf_counter()
        z = m.decode_device(out.streams, out.lens, n, 64, 96, lanes=0); torch.cuda.synchronize()
        dd = time.perf_counter() - t
        print(f"n={n} ws={ws}: encode {dt*1e3:.1f} ms {n*512*768/dt/1e6:.1f} Mpix/s | decode {dd*1e3:.1f} ms {n*512*768/dd/1e6:.1f} Mpix/s  ok={bool(torch.equal(z, out.zhat))}", flush=True)
    del x, out, z
PY
grep -E "^exit|passed|failed|^==|Error|^n=" gpurun_out/check9.log
