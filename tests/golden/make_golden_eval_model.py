"""Runs the UNMODIFIED reference agent's eval_model (agents/blkbsdimgcomp_agent.py:560-641) -- config JSON, checkpoint
file, PNG folder, dataloader, compress, decompress, metrics, log line -- on ONE synthetic image with the synthetic
weights, on the CPU (the reference's own setting there: one thread, AGENT:565-566), and stores what its log line
reports.  The GPU test test_eval_model_matches_reference_log_line reproduces the same image through the B200 path.

    python tests/golden/make_golden_eval_model.py        (~1 min)
Writes tests/golden/eval_model_B8_lowrate_208x176.json.
"""
import json
import logging
import os
import re
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import lbic_b200  # noqa: E402
from lbic_b200 import weights  # noqa: E402
from oracle.ref_shim import load_reference  # noqa: E402

H, W, IMG_SEED, WEIGHT_SEED = 176, 208, 77, 1337
LINE = re.compile(r"Image\s+(\d+) --> RDLoss:([-\d.eE]+) MSE/PSNR:([-\d.eE]+)/([-\d.eE]+) Rate:([-\d.eE]+) "
                  r"MS-SSIM/dB:([-\d.eE]+)/([-\d.eE]+) Enc/DecTime:([-\d.eE]+)/([-\d.eE]+) "
                  r"Enc-Dec.Mad/Max/Min:([-\d.eEna]+)/([-\d.eEna]+)/([-\d.eEna]+) \((.*)\)")


def reference_config(tmp, backend=None):
    """The reference's own configs/blkbsdimgcomp_B8_lowrate.json with the paths pointed at `tmp` and cuda off."""
    cfg = json.load(open(os.path.join(load_reference.REF, "configs", "blkbsdimgcomp_B8_lowrate.json")))
    cfg.update(cuda=False, mode="eval_model", lambda_=117.045, exp_name="exp_117.045",
               valid_data=os.path.join(tmp, "kodak", "test"), test_data=os.path.join(tmp, "kodak", "test"),
               modelbest_file_load="model_best.pth.tar",
               checkpoint_dir=os.path.join(tmp, "exp", "checkpoints") + "/", out_dir=os.path.join(tmp, "exp", "out") + "/",
               log_dir=os.path.join(tmp, "exp", "logs") + "/", summary_dir=os.path.join(tmp, "exp", "summaries") + "/")
    for k in ("train_data_1", "train_data_2", "train_data_3", "train_data_4"):
        cfg[k] = cfg["valid_data"]
    cfg["num_train_dirs"] = 1
    if backend:
        cfg["backend"] = backend
    return load_reference.EasyDict(cfg)


def prepare_inputs(tmp):
    """One PNG in the validation folder + a weights-only checkpoint (experiments/extract_model_weights_only.py layout)."""
    from PIL import Image
    cfg = lbic_b200.load_config("B8_lowrate")
    os.makedirs(os.path.join(tmp, "kodak", "test"))
    for d in ("checkpoints", "out", "logs", "summaries"):
        os.makedirs(os.path.join(tmp, "exp", d))
    img = weights.synth_image_u8(H, W, IMG_SEED)
    Image.fromarray(np.ascontiguousarray(img.transpose(1, 2, 0)), "RGB").save(os.path.join(tmp, "kodak", "test", "synth01.png"))
    sd = weights.synth_state_dict(cfg, WEIGHT_SEED)
    torch.save({"state_dict0": sd}, os.path.join(tmp, "exp", "checkpoints", "model_best.pth.tar"))
    return img


def run_eval_model(agent_cls, config):
    """agent_cls(config).eval_model() with the 'Agent' logger captured -> parsed per-image records."""
    records, lines = [], []

    class Grab(logging.Handler):
        def emit(self, rec):
            lines.append(rec.getMessage())

    lg = logging.getLogger("Agent")
    lg.setLevel(logging.INFO)
    h = Grab()
    lg.addHandler(h)
    exact = []
    try:
        agent = agent_cls(config)
        stock_compress = agent.model0.compress

        def compress(x, LRU, chlat):                 # observe only: exact stream size and reconstruction error
            bitstream, zhat = stock_compress(x, LRU, chlat)
            exact.append(dict(bytes=len(bitstream), mse_padded=float(((x - zhat) ** 2).mean())))
            return bitstream, zhat

        agent.model0.compress = compress
        agent.eval_model()
    finally:
        lg.removeHandler(h)
    for ln in lines:
        m = LINE.search(ln)
        if m:
            g = m.groups()
            records.append(dict(image=int(g[0]), rd_loss=float(g[1]), mse=float(g[2]), psnr=float(g[3]), bpp=float(g[4]),
                                msssim=float(g[5]), msssim_db=float(g[6]), enc_s=float(g[7]), dec_s=float(g[8]),
                                enc_dec_mad=float(g[9]), enc_dec_max=float(g[10]), enc_dec_min=float(g[11]), file=g[12]))
    for r, e in zip(records, exact):
        r.update(e)
    return records, lines


def main():
    mod = load_reference.load_agent()
    with tempfile.TemporaryDirectory() as tmp:
        prepare_inputs(tmp)
        recs, lines = run_eval_model(mod.BlockBasedImgCompLossyAgent, reference_config(tmp))
        assert len(recs) == 1, lines
        r = recs[0]
        out = dict(_about="what the UNMODIFIED reference agent's eval_model logs for one synthetic image on the CPU; "
                          "MS-SSIM is computed by a stand-in (pytorch_msssim is not installed) and is not a pin",
                   config="B8_lowrate", H=H, W=W, image_seed=IMG_SEED, weight_seed=WEIGHT_SEED, record=r,
                   log_line=[ln for ln in lines if ln.startswith("Image")][0])
        json.dump(out, open(os.path.join(HERE, f"eval_model_B8_lowrate_{W}x{H}.json"), "w"), indent=1)
        print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
