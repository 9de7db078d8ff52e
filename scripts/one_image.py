"""compress + decompress of ONE 768x512 image (the reference's call pattern), for profiling the wave kernel."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lbic_b200
from lbic_b200 import weights
from lbic_b200.layout import arrange_block_pixels_to_channel_dim
from lbic_b200.net import BlockBasedImgCompLossyNetv9
dev = torch.device("cuda:0")
cfg = lbic_b200.load_config("B8_lowrate")
m = BlockBasedImgCompLossyNetv9(cfg, device=dev)
m.load_state_dict(weights.synth_state_dict(cfg, 1337)); m.update(force=True)
x = arrange_block_pixels_to_channel_dim((weights.synth_images(1, 512, 768) - 0.5).to(dev), 8)
for _ in range(2):
    s, z = m.compress(x, [1, 1, 1], 96)
    zd = m.decompress(s, [1, 1, 1], x.shape, 96, dev)
torch.cuda.synchronize()
print("identical", bool(torch.equal(z, zd)), len(s))
