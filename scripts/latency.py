"""Latency of the reference's real call pattern: compress / decompress of ONE image (and small batches, large images),
device-resident, CUDA events, best of 3 after a warm-up; the persistent wavefront kernel (gemm_wave.cu) on and off.
    python scripts/latency.py [--quick] > gpurun_out/latency.jsonl
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lbic_b200  # noqa: E402
from lbic_b200 import weights  # noqa: E402
from lbic_b200.layout import arrange_block_pixels_to_channel_dim  # noqa: E402
from lbic_b200.net import BlockBasedImgCompLossyNetv9  # noqa: E402

dev = torch.device("cuda:0")


def timed(fn, reps=3):
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        out = fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best, out


def run(cfgname, n, H, W, lanes_list=(0, 1), waves=(0, 1), max_ms_for_ref=1e9):
    cfg = lbic_b200.load_config(cfgname)
    m = BlockBasedImgCompLossyNetv9(cfg, device=dev)
    m.load_state_dict(weights.synth_state_dict(cfg, 1337))
    m.update(force=True)
    B = int(cfg.block_size)
    g = torch.Generator(device=dev)
    g.manual_seed(5)
    x = arrange_block_pixels_to_channel_dim(torch.rand(n, 3, H, W, device=dev, generator=g) - 0.5, B)
    ref = {}
    for lanes in lanes_list:
        for wave in waves:
            m.set_option("wave", wave)
            if wave:
                m.set_option("wave_max_rows", 4096)
            out = m.encode_device(x, lanes=lanes)
            te, out = timed(lambda: m.encode_device(x, lanes=lanes, out=out))
            Hb, Wb = H // B, W // B
            z = m.decode_device(out.streams, out.lens, n, Hb, Wb, lanes=lanes)
            td, z = timed(lambda: m.decode_device(out.streams, out.lens, n, Hb, Wb, lanes=lanes))
            lens = out.lens.cpu()
            key = lanes
            same = None
            if key in ref:
                same = bool(torch.equal(ref[key][0], out.zhat)) and bool(torch.equal(ref[key][1], lens))
            else:
                ref[key] = (out.zhat.clone(), lens)
            print(json.dumps(dict(config=cfgname, images=n, W=W, H=H, container="reference" if lanes == 1 else "lane",
                                  wave_kernel=bool(wave), encode_ms=round(te, 3), decode_ms=round(td, 3),
                                  encode_mpix_s=round(n * H * W / te / 1e3, 2), decode_mpix_s=round(n * H * W / td / 1e3, 2),
                                  enc_dec_identical=bool(torch.equal(z, out.zhat)), identical_to_other_path=same)), flush=True)
    del m


if __name__ == "__main__":
    quick = "--quick" in sys.argv
    run("B8_lowrate", 1, 512, 768)
    if not quick:
        for n in (8, 24, 64):
            run("B8_lowrate", n, 512, 768, lanes_list=(0,))
        run("B16_lowrate", 1, 2048, 2048, lanes_list=(0,))
        run("B16_lowrate", 8, 2048, 2048, lanes_list=(0,))
        run("B8_lowrate", 1, 4096, 4096, lanes_list=(0,))
        run("B8_lowrate", 1, 8192, 8192, lanes_list=(0,))
