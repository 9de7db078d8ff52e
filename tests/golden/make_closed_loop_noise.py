"""Does ONE rounding flip cascade through the reference's closed loop?  

Runs the UNMODIFIED reference `compress` (graphs/models/BlockBasedImgCompLossy_net.py:319-377, imported through
oracle/ref_shim) on the full-size golden image (B8 KS3111 N768 M96, 768x512, 589 824 symbols) in arithmetic that is
mathematically the same network but not bit-identical to the stock fp32 run:

  perm<k>  fp32, the hidden channels of every layer permuted consistently (weights, biases, masks, GDN beta/gamma):
           the same real-valued function, the same reference code, the same fp32 torch ops -- only the ORDER in which
           each dot product is accumulated differs;
  fp64     the reference model after .double() (the exact-arithmetic answer to ~1e-16);
  flip:v:h:c   the stock fp32 run with ONE symbol forced to the other neighbouring integer at block (v,h), channel c
           (the only intervention: `quantize` returns the other rounding for that one element) -- the cleanest
           measurement of how far a single flip propagates through the reference's own loop.  flip:9:53:53 is the
           position where the B200 path's first flip occurred in round 1 (profiles/r1_full_size_parity.json).

and counts closed-loop symbol mismatches against the stock fp32 run stored in full_B8_lowrate_768x512.npz.  If a
single boundary flip de-synchronises the rest of the image here too, whole-image closed-loop symbol identity is not a
property any non-bit-identical implementation (GPU, another BLAS, another thread count) can have; the tests then
assert identity up to the first flip + that every teacher-forced mismatch is a rounding-boundary case.

    python tests/golden/make_closed_loop_noise.py [perm1 perm2 fp64 ...]     (about 2-4 min per variant on 8 threads)
Writes tests/golden/closed_loop_noise_B8_lowrate_768x512.json.
"""
import json
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

OUT = os.path.join(HERE, "closed_loop_noise_B8_lowrate_768x512.json")


def permuted_state_dict(sd, cfg, seed):
    """The same network with every hidden channel axis permuted (outputs y, ksi, xhat and inputs x, zhat keep their
    order).  Purely a re-labelling: in exact arithmetic the model computes the identical function."""
    w = weights.widths(cfg)
    g = torch.Generator()
    g.manual_seed(seed)
    perm = {n: torch.randperm(w[n], generator=g) for n in ("N", "C2", "C3", "E1", "E2", "E3")}
    fN, fC2, fC3 = (torch.randperm(w[n], generator=g) for n in ("N", "C2", "C3"))   # encoder chain
    iN, iC2, iC3 = perm["N"], perm["C2"], perm["C3"]                                   # decoder chain
    out = {k: v.clone() for k, v in sd.items()}

    def conv(prefix, p_out, p_in):
        for s in (".weight", ".mask"):
            t = out[prefix + s]
            if p_out is not None:
                t = t[p_out]
            if p_in is not None:
                t = t[:, p_in]
            out[prefix + s] = t.contiguous()
        if p_out is not None:
            out[prefix + ".bias"] = out[prefix + ".bias"][p_out].contiguous()

    def gdn(prefix, p):
        out[prefix + ".beta"] = out[prefix + ".beta"][p].contiguous()
        out[prefix + ".gamma"] = out[prefix + ".gamma"][p][:, p].contiguous()

    conv("prtr_forward1", fN, None); conv("prtr_forward2", fN, None)
    gdn("prtr_forward3.0", fN); conv("prtr_forward3.1", fC2, fN)
    gdn("prtr_forward3.2", fC2); conv("prtr_forward3.3", fC3, fC2)
    gdn("prtr_forward3.4", fC3); conv("prtr_forward3.5", None, fC3)
    conv("prtr_inverse1", iN, None); conv("prtr_inverse2", iN, None)
    gdn("prtr_inverse3.0", iN); conv("prtr_inverse3.1", iC2, iN)
    gdn("prtr_inverse3.2", iC2); conv("prtr_inverse3.3", iC3, iC2)
    gdn("prtr_inverse3.4", iC3); conv("prtr_inverse3.5", None, iC3)
    conv("get_meanscale.0", perm["E1"], None); conv("get_meanscale.2", perm["E2"], perm["E1"])
    conv("get_meanscale.4", perm["E3"], perm["E2"]); conv("get_meanscale.6", None, perm["E3"])
    return out


def run_variant(ref, cfg, sd, x, dtype, flip=None):
    import compressai.ans as ans
    captured = {}
    orig_flush = ans.BufferedRansEncoder.flush

    def flush(self):
        captured["symbols"] = np.asarray(self._symbols, dtype=np.int32)
        captured["indexes"] = np.asarray(self._indexes, dtype=np.int32)
        return orig_flush(self)

    ans.BufferedRansEncoder.flush = flush
    try:
        m = ref.BlockBasedImgCompLossyNetv9(cfg).eval()
        m.load_state_dict(sd, strict=False)
        m.update(force=True)
        if dtype == torch.float64:
            m = m.double()
        if flip is not None:
            fv, fh, fc = flip
            wd = x.shape[3]
            g = m.conditional_gaussian_model
            stock_quantize, calls = g.quantize, [0]

            def quantize(inputs, mode, means=None):
                out = stock_quantize(inputs, mode, means)
                if calls[0] == fv * wd + fh:
                    d = float((inputs - means)[0, fc, 0, 0])
                    lo = int(np.floor(d))
                    other = lo + 1 if int(out[0, fc, 0, 0]) == lo else lo
                    captured["forced"] = dict(y_minus_mean=d, stock=int(out[0, fc, 0, 0]), forced=other)
                    out = out.clone()
                    out[0, fc, 0, 0] = other
                calls[0] += 1
                return out

            g.quantize = quantize
        t0 = time.time()
        with torch.no_grad():
            stream, zhat = m.compress(x.to(dtype), [1, 1, 1], cfg.M)
        secs = time.time() - t0
    finally:
        ans.BufferedRansEncoder.flush = orig_flush
    return captured["symbols"], captured["indexes"], stream, zhat.float(), secs, captured.get("forced")


def main():
    variants = sys.argv[1:] or ["perm1", "perm2", "fp64"]
    torch.set_num_threads(8)
    torch.use_deterministic_algorithms(True)
    ref = load_reference.load()
    cfg = lbic_b200.load_config("B8_lowrate")
    gold = np.load(os.path.join(HERE, "full_B8_lowrate_768x512.npz"))
    H, W = int(gold["H"]), int(gold["W"])
    Hb, Wb, M = H // 8, W // 8, int(cfg.M)
    sd = weights.synth_state_dict(cfg, int(gold["seed"]))
    x_img = weights.u8_to_model_input(weights.synth_image_u8(H, W, int(gold["image_seed"])))
    x = nets.arrange_block_pixels_to_channel_dim(x_img, 8)
    ref_sym = gold["symbols"].astype(np.int32).reshape(-1)
    ref_idx = gold["indexes"].astype(np.int32).reshape(-1)
    results = json.load(open(OUT)) if os.path.exists(OUT) else {}
    results["_about"] = ("closed-loop symbol mismatches of the UNMODIFIED reference compress() against its own stock fp32 run "
                         "(full_B8_lowrate_768x512.npz) when only the accumulation order (perm*) or the precision (fp64) "
                         "changes; generated by tests/golden/make_closed_loop_noise.py")
    results["symbols"] = int(ref_sym.size)
    P = nets.effective_params(sd, cfg)
    for name in variants:
        if name.startswith("perm"):
            sdv, dt = permuted_state_dict(sd, cfg, 100 + int(name[4:] or 1)), torch.float32
        elif name == "fp64":
            sdv, dt = sd, torch.float64
        elif name == "stock":
            sdv, dt = sd, torch.float32
        elif name.startswith("flip:"):
            sdv, dt = sd, torch.float32
        else:
            raise SystemExit(f"unknown variant {name}")
        flip = tuple(int(t) for t in name.split(":")[1:]) if name.startswith("flip:") else None
        sym, idx, stream, zhat, secs, forced = run_variant(ref, cfg, sdv, x, dt, flip)
        mis = sym != ref_sym
        n_mis = int(mis.sum())
        rec = dict(symbol_mismatches=n_mis, index_mismatches=int((idx != ref_idx).sum()),
                   mismatch_fraction=n_mis / ref_sym.size, stream_bytes=len(stream), ref_stream_bytes=int(gold["stream_len"]),
                   seconds=round(secs, 1))
        if forced:
            rec["forced"] = forced
        if n_mis:
            first = int(np.argmax(mis))
            blk, ch = divmod(first, M)
            v, h = divmod(blk, Wb)
            rec["first_mismatch"] = dict(block=[v, h], channel=ch, ref_symbol=int(ref_sym[first]), symbol=int(sym[first]))
            # distance of the stock fp32 model's y - mean from the rounding boundary at that position, teacher-forced on
            # this variant's own reconstruction (all inputs of that block are identical in both runs up to the first flip)
            s_tf, i_tf, _, y, ksi = nets.whole_image_eval(P, x, zhat)
            d = (y - ksi[:, M:])[0, ch, v, h].item()
            rec["first_mismatch"]["y_minus_mean_fp32"] = d
            rec["first_mismatch"]["boundary_distance"] = abs(abs(d - np.floor(d)) - 0.5)
            rec["blocks_before_first_mismatch"] = blk
            rec["mismatches_after_first_fraction"] = n_mis / max(1, ref_sym.size - first)
            # teacher-forced: the stock fp32 nets on this variant's final zhat vs this variant's symbols
            rec["teacher_forced_symbol_mismatches"] = int((s_tf[0].reshape(-1).numpy() != sym).sum())
        results[name] = rec
        print(name, json.dumps(rec), flush=True)
        json.dump(results, open(OUT, "w"), indent=1)


if __name__ == "__main__":
    main()
