"""Full-size golden: the UNMODIFIED reference's compress() on ONE 768x512 image (default B8 KS3111 N768 M96, grid
64x96, 589 824 symbols: ~1-2 minutes on 8 CPU threads; B8_highrate / B4_highrate (KS3311): 10-30 minutes).  Stores
symbols (int8), indexes (uint8), the bitstream's length + sha256, and the reconstruction PSNR; the input is regenerated
bit-exactly by weights.synth_image_u8.
    python tests/golden/make_golden_full.py [--config B8_highrate] [--height 512 --width 768]
"""
import argparse
import hashlib
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import lbic_b200  # noqa: E402
from lbic_b200 import weights  # noqa: E402
from oracle import nets  # noqa: E402
from oracle.ref_shim import load_reference  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="B8_lowrate")
    ap.add_argument("--height", type=int, default=512)
    ap.add_argument("--width", type=int, default=768)
    args = ap.parse_args()
    torch.set_num_threads(8)
    torch.use_deterministic_algorithms(True)
    ref = load_reference.load()
    cfg = lbic_b200.load_config(args.config)
    sd = weights.synth_state_dict(cfg, 1337)
    m = ref.BlockBasedImgCompLossyNetv9(cfg).eval()
    m.load_state_dict(sd, strict=False)
    m.update(force=True)
    H, W, seed = args.height, args.width, 2024
    B = int(cfg.block_size)
    x_img = weights.u8_to_model_input(weights.synth_image_u8(H, W, seed))
    x = nets.arrange_block_pixels_to_channel_dim(x_img, B)
    r = sum(int(k) // 2 for k in cfg.KS)
    t0 = time.time()
    with torch.no_grad():
        stream, zhat = m.compress(x, [r, r, r], cfg.M)
    t_ref = time.time() - t0
    # symbols / indexes: the oracle loop (bit-identical to the reference on every small golden case) must reproduce
    # the reference's bitstream here too; its symbols are then the reference's symbols.
    P = nets.effective_params(sd, cfg)
    g = m.conditional_gaussian_model
    tabs = (g.quantized_cdf, g.cdf_length, g.offset)
    ostream, ozhat, s_loop, i_loop = nets.compress(P, tabs, x)
    assert ostream == stream and torch.equal(ozhat, zhat), "oracle loop differs from the reference at full size"
    s, i = s_loop.unsqueeze(0), i_loop.unsqueeze(0)
    # fp32 noise floor: one batched fp32 evaluation on the reference's own zhat (different conv accumulation order)
    s_fp, i_fp, _, _, _ = nets.whole_image_eval(P, x, zhat)
    fp32_sym_mis = int((s_fp[0] != s_loop).sum())
    fp32_idx_mis = int((i_fp[0] != i_loop).sum())
    print(f"fp32 batched-vs-per-block conv (CPU, same weights, same zhat): {fp32_sym_mis} symbol / {fp32_idx_mis} index "
          f"mismatches of {s_loop.numel()}")
    mse = float(((x - zhat) ** 2).mean())
    np.savez_compressed(os.path.join(HERE, f"full_{args.config}_{W}x{H}.npz"), config=args.config, seed=1337,
                        H=H, W=W, image_seed=seed, symbols=s[0].numpy().astype(np.int16 if int(s.abs().max()) > 127 else np.int8),
                        indexes=i[0].numpy().astype(np.uint8), stream_len=len(stream),
                        stream_sha256=hashlib.sha256(stream).hexdigest(), psnr=-10.0 * np.log10(mse),
                        zhat_sha256=hashlib.sha256(zhat.numpy().tobytes()).hexdigest(), ref_seconds=t_ref,
                        fp32_noise_symbol_mismatches=fp32_sym_mis, fp32_noise_index_mismatches=fp32_idx_mis)
    print(f"reference compress: {t_ref:.1f} s, {len(stream)} bytes, psnr {-10 * np.log10(mse):.3f} dB, "
          f"|sym|max {int(s.abs().max())}")


if __name__ == "__main__":
    main()
