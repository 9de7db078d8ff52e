import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lbic_b200
from lbic_b200.net import BlockBasedImgCompLossyNetv9
from lbic_b200.weights import synth_state_dict, synth_images
from lbic_b200.layout import arrange_block_pixels_to_channel_dim
cfg = lbic_b200.load_config("B8_lowrate")
m = BlockBasedImgCompLossyNetv9(cfg, device="cuda:0")
m.load_state_dict(synth_state_dict(cfg)); m.update()
n = 37
img = synth_images(n, 7 * 8, 12 * 8, seed0=57)
x = arrange_block_pixels_to_channel_dim((img - 0.5).cuda(), 8)
m.set_option("flow", 0)
r = m.encode_device(x, lanes=0, want_symbols=True)
zr = m.decode_device(r.streams, r.lens, n, 7, 12, lanes=0)
m.set_option("flow", 2)
zd, sd = m.decode_device(r.streams, r.lens, n, 7, 12, lanes=0, want_symbols=True)
print("LBIC_FLOW_PART", os.environ.get("LBIC_FLOW_PART"), "decode sym mismatches", int((sd != r.sym).sum()), "zhat maxdiff", float((zd - zr).abs().max()))
