import json, os, sys, torch
sys.path.insert(0, os.getcwd())
import lbic_b200
from lbic_b200 import weights
from lbic_b200.layout import arrange_block_pixels_to_channel_dim
from lbic_b200.net import BlockBasedImgCompLossyNetv9
dev = torch.device("cuda:0")
def timed(fn, reps=3):
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record(); out = fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best, out
for cfgname, n in (("B8_lowrate", 1), ("B8_lowrate", 2), ("B4_highrate", 1), ("B8_highrate", 2), ("B16_lowrate", 4)):
    cfg = lbic_b200.load_config(cfgname)
    m = BlockBasedImgCompLossyNetv9(cfg, device=dev)
    m.load_state_dict(weights.synth_state_dict(cfg, 1337)); m.update(force=True)
    B = int(cfg.block_size)
    x = arrange_block_pixels_to_channel_dim(torch.rand(n, 3, 512, 768, device=dev) - 0.5, B)
    out = m.encode_device(x, lanes=0)
    for cap in (64, 128):
        m.set_option("wave_dec_max_rows", cap)
        z = m.decode_device(out.streams, out.lens, n, 512 // B, 768 // B, lanes=0)
        td, z = timed(lambda: m.decode_device(out.streams, out.lens, n, 512 // B, 768 // B, lanes=0))
        print(json.dumps(dict(config=cfgname, images=n, wave_dec_max_rows=cap, decode_ms=round(td, 2), identical=bool(torch.equal(z, out.zhat)))), flush=True)
    del m
