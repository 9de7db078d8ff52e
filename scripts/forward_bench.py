"""Throughput of the open-loop forward (lbic_forward) on 256x256 patches, the ACL training-set regeneration workload
(AGENT:643-684: ~19 k patches per ACL iteration, 19 min in the reference's logs)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lbic_b200
from lbic_b200.net import BlockBasedImgCompLossyNetv9
from lbic_b200.weights import synth_state_dict, synth_images
from lbic_b200.layout import arrange_block_pixels_to_channel_dim
name = sys.argv[1] if len(sys.argv) > 1 else "B8_lowrate"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
cfg = lbic_b200.load_config(name)
net = BlockBasedImgCompLossyNetv9(cfg, device="cuda:0")
net.load_state_dict(synth_state_dict(cfg)); net.update()
B = int(cfg.block_size)
img = (synth_images(8, 256, 256) - 0.5).repeat((n + 7) // 8, 1, 1, 1)[:n].cuda()
x = arrange_block_pixels_to_channel_dim(img, B).contiguous(); del img
zhat = (x + 0.02 * torch.randn_like(x)).clamp_(-0.5, 0.5)
net.forward(zhat, x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    xhat, info = net.forward(zhat, x, clamp=True)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(f"{name}: {n} patches of 256x256 in {ms:.1f} ms = {n * 65536 / ms / 1e3:.0f} Mpixel/s "
      f"({19000 * ms / n / 1e3:.2f} s per 19 k-patch ACL regeneration; bits/pixel {float(info.sum()) / (n * 65536):.2f})")
