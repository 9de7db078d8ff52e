"""One 768x512 image through encode_device (and decode_device with `dec`) of one topology: the workload of the wave-kernel
captures.  LBIC_TRACE_CONFIG picks the topology (default B8_lowrate).  python scripts/one_image_wave.py enc|dec"""
import os, sys, torch
sys.path.insert(0, os.getcwd())
import lbic_b200
from lbic_b200 import weights
from lbic_b200.layout import arrange_block_pixels_to_channel_dim
from lbic_b200.net import BlockBasedImgCompLossyNetv9
dev = torch.device("cuda:0")
cfg = lbic_b200.load_config(os.environ.get("LBIC_TRACE_CONFIG", "B8_lowrate"))
m = BlockBasedImgCompLossyNetv9(cfg, device=dev)
m.load_state_dict(weights.synth_state_dict(cfg, 1337)); m.update(force=True)
B = int(cfg.block_size)
x = arrange_block_pixels_to_channel_dim(torch.rand(1, 3, 512, 768, device=dev) - 0.5, B)
o = m.encode_device(x, lanes=0)
torch.cuda.synchronize()
if len(sys.argv) > 1 and sys.argv[1] == "dec":
    m.decode_device(o.streams, o.lens, 1, 512 // B, 768 // B, lanes=0)
    torch.cuda.synchronize()
print("done")
