"""Per-layer GEMM time / throughput inside a real encode + decode (CUDA events around every launch)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lbic_b200
from lbic_b200.net import BlockBasedImgCompLossyNetv9
from lbic_b200.weights import synth_state_dict, synth_images

name = sys.argv[1] if len(sys.argv) > 1 else "B8_lowrate"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
H = int(sys.argv[3]) if len(sys.argv) > 3 else 512
W = int(sys.argv[4]) if len(sys.argv) > 4 else 768
cfg = lbic_b200.load_config(name)
net = BlockBasedImgCompLossyNetv9(cfg, device="cuda:0")
net.load_state_dict(synth_state_dict(cfg))
net.update()
from lbic_b200.layout import arrange_block_pixels_to_channel_dim
img = (synth_images(8, H, W) - 0.5).repeat((n + 7) // 8, 1, 1, 1)[:n].cuda()
x = arrange_block_pixels_to_channel_dim(img, int(cfg.block_size)).contiguous()
del img
Hb, Wb = x.shape[2], x.shape[3]
enc = net.encode_device(x, lanes=0)          # warm-up
net.decode_device(enc.streams, enc.lens, n, Hb, Wb, lanes=0)
for what in ("encode", "decode"):
    net.set_profiling(True)
    if what == "encode":
        enc = net.encode_device(x, lanes=0)
    else:
        net.decode_device(enc.streams, enc.lens, n, Hb, Wb, lanes=0)
    prof = net.get_layer_profile()
    net.set_profiling(False)
    tot = sum(v["ms"] for v in prof.values())
    print(f"# {name} n={n} {H}x{W} {what}: GEMM total {tot:.1f} ms")
    for k, v in prof.items():
        if v["launches"]:
            tf = v["flops"] / (v["ms"] * 1e-3) / 1e12
            print(f"  {k:4s} launches {v['launches']:5d}  {v['ms']:8.2f} ms  {100*v['ms']/tot:5.1f}%  {tf:6.1f} TF/s alg ({3*tf:6.0f} mma)")
