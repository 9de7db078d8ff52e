"""Decode of a 1024-image batch under cudaProfilerStart/Stop (for an ncu launch list of the decode path only)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lbic_b200
from lbic_b200.net import BlockBasedImgCompLossyNetv9
from lbic_b200.weights import synth_state_dict, synth_images
from lbic_b200.layout import arrange_block_pixels_to_channel_dim
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
lanes = int(sys.argv[2]) if len(sys.argv) > 2 else 0
cfg = lbic_b200.load_config("B8_lowrate")
net = BlockBasedImgCompLossyNetv9(cfg, device="cuda:0")
net.load_state_dict(synth_state_dict(cfg)); net.update()
img = (synth_images(8, 512, 768) - 0.5).repeat((n + 7) // 8, 1, 1, 1)[:n].cuda()
x = arrange_block_pixels_to_channel_dim(img, 8).contiguous(); del img
Hb, Wb = x.shape[2], x.shape[3]
enc = net.encode_device(x, lanes=lanes)
net.decode_device(enc.streams, enc.lens, n, Hb, Wb, lanes=lanes)
torch.cuda.synchronize()
torch.cuda.profiler.start()
if len(sys.argv) > 3 and sys.argv[3] == "enc":
    net.encode_device(x, lanes=lanes)
else:
    net.decode_device(enc.streams, enc.lens, n, Hb, Wb, lanes=lanes)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
