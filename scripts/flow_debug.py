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
n = int(sys.argv[1]) if len(sys.argv) > 1 else 37
img = synth_images(n, 7 * 8, 12 * 8, seed0=57)
x = arrange_block_pixels_to_channel_dim((img - 0.5).cuda(), 8)
m.set_option("flow", 0)
r = m.encode_device(x, lanes=0, want_symbols=True)
zr = m.decode_device(r.streams, r.lens, n, 7, 12, lanes=0)
print("ref enc/dec identical", torch.equal(zr, r.zhat))
m.set_option("flow", 2)
zd = m.decode_device(r.streams, r.lens, n, 7, 12, lanes=0)
print("flow DECODE == ref:", torch.equal(zd, zr), float((zd - zr).abs().max()))
g = m.encode_device(x, lanes=0, want_symbols=True)
ds = (g.sym != r.sym)
print("flow ENCODE sym mismatches", int(ds.sum()), "of", ds.numel(), "idx mismatches", int((g.idx != r.idx).sum()), "zhat maxdiff", float((g.zhat - r.zhat).abs().max()))
if ds.any():
    bad = ds.any(dim=3)          # (n, Hb, Wb)
    print("per image bad blocks:", bad.flatten(1).sum(1)[:10].tolist())
    i = int(bad.flatten(1).any(1).nonzero()[0])
    print("image", i, "bad map:\n", bad[i].int())
    v, h = [int(t) for t in bad[i].nonzero()[0]]
    print("first bad block", v, h, "ref", r.sym[i, v, h, :8].tolist(), "got", g.sym[i, v, h, :8].tolist())
