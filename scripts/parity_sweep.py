"""Full-size parity statistics for every reference topology: for each of >= 8 images per topology
the B200 closed-loop encode is checked against ONE oracle evaluation (torch CPU fp32, oracle/nets.py whole_image_eval)
on the GPU's own reconstruction -- the closed loop's fixed point, SURVEY.md fact 10 -- and every teacher-forced mismatch
is measured against its rounding boundary.  One JSON line per image:

    config, image, symbols, tf_symbol_mismatches (ppm), tf_index_mismatches, worst boundary distances, first flip block,
    zhat max difference on blocks with identical symbols, bytes, enc/dec identical

    python scripts/parity_sweep.py [--images 8] > gpurun_out/parity_sweep.jsonl      (on the GPU box)
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import lbic_b200  # noqa: E402
from lbic_b200 import weights  # noqa: E402
from lbic_b200.layout import arrange_block_pixels_to_channel_dim  # noqa: E402
from lbic_b200.net import BlockBasedImgCompLossyNetv9  # noqa: E402
from oracle import nets  # noqa: E402
from test_gpu_parity import tf_boundary_report  # noqa: E402

WORK = [("B8_lowrate", 512, 768, 8), ("B8_highrate", 512, 768, 8), ("B4_highrate", 512, 768, 8),
        ("B16_lowrate", 2048, 2048, 4), ("B16_lowrate", 512, 768, 8)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=0, help="override the number of images per topology")
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    for cfgname, H, W, n in WORK:
        if args.only and args.only != cfgname:
            continue
        n = args.images or n
        cfg = lbic_b200.load_config(cfgname)
        sd = weights.synth_state_dict(cfg, 1337)
        m = BlockBasedImgCompLossyNetv9(cfg, device=dev)
        m.load_state_dict(sd)
        m.update(force=True)
        P = nets.effective_params(sd, cfg)
        B = m.B
        imgs = torch.stack([weights.u8_to_model_input(weights.synth_image_u8(H, W, 3000 + i))[0] for i in range(n)])
        x = arrange_block_pixels_to_channel_dim(imgs.to(dev), B)
        strings, zhat, sym, idx = m.compress_batch(x, lanes=0, return_symbols=True)
        zdec = m.decompress_batch(strings, x.shape, lanes=0)
        same = bool(torch.equal(zdec, zhat))
        for i in range(n):
            t0 = time.time()
            xc, zc = x[i:i + 1].cpu(), zhat[i:i + 1].cpu()
            sc, ic = sym[i:i + 1].cpu(), idx[i:i + 1].cpu().int()
            s2, i2, xh2, y2, ksi2 = nets.whole_image_eval(P, xc, zc)
            rep = tf_boundary_report(s2, i2, y2, ksi2, sc, ic, P.scale_table)
            mis = (s2 != sc)
            first = None
            if bool(mis.any()):
                blk = int(mis[0].any(dim=-1).reshape(-1).float().argmax())
                first = [blk // sc.shape[2], blk % sc.shape[2]]
            same_blk = (~mis.any(dim=-1)).unsqueeze(1)
            rec = dict(config=cfgname, H=H, W=W, image=i, symbols=sc.numel(), tf_symbol_mismatches=int(mis.sum()),
                       tf_symbol_ppm=1e6 * int(mis.sum()) / sc.numel(), tf_index_mismatches=int((i2 != ic).sum()),
                       first_tf_flip_block=first, zhat_maxdiff_same_blocks=float(((xh2 - zc).abs() * same_blk).max()),
                       sym_std=float(sc.float().std()), sym_absmax=int(sc.abs().max()), bytes=len(strings[i]),
                       bpp=8.0 * len(strings[i]) / (H * W), enc_dec_identical=same, oracle_seconds=round(time.time() - t0, 1),
                       **rep)
            print(json.dumps(rec), flush=True)
        del m


if __name__ == "__main__":
    main()
