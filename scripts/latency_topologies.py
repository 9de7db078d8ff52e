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
# LBIC_LAT_CONFIGS=a,b restricts the topologies, LBIC_LAT_LANE_ONLY=1 skips the (raster-serial) reference container
CONFIGS = os.environ.get("LBIC_LAT_CONFIGS", "B8_lowrate,B8_highrate,B4_highrate,B16_lowrate").split(",")
LANES = (0,) if os.environ.get("LBIC_LAT_LANE_ONLY") == "1" else (0, 1)
for cfgname in CONFIGS:
    cfg = lbic_b200.load_config(cfgname)
    m = BlockBasedImgCompLossyNetv9(cfg, device=dev)
    m.load_state_dict(weights.synth_state_dict(cfg, 1337)); m.update(force=True)
    B = int(cfg.block_size)
    x = arrange_block_pixels_to_channel_dim(torch.rand(1, 3, 512, 768, device=dev) - 0.5, B)
    for lanes in LANES:
        out = m.encode_device(x, lanes=lanes)
        l0 = m.launch_count()
        te, out = timed(lambda: m.encode_device(x, lanes=lanes, out=out))
        le = (m.launch_count() - l0) // 3
        z = m.decode_device(out.streams, out.lens, 1, 512 // B, 768 // B, lanes=lanes)
        td, z = timed(lambda: m.decode_device(out.streams, out.lens, 1, 512 // B, 768 // B, lanes=lanes))
        print(json.dumps(dict(config=cfgname, KS="".join(map(str, cfg.KS)), container="reference" if lanes else "lane", encode_ms=round(te, 2), decode_ms=round(td, 2), launches_per_encode=le, identical=bool(torch.equal(z, out.zhat)))), flush=True)
    del m
