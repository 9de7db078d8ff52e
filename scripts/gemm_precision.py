"""Where does the B200 path's rounding noise come from?  

D = A W^T for layer-like shapes through
  tc    the product GEMM core: fp16 hi/lo planes, 3 tcgen05.mma passes, fp32 TMEM accumulator
  simt  the same hi/lo planes multiplied and accumulated with fp32 FFMA (the cross-check twin)
  f32   torch fp32 matmul on the GPU (cuBLAS SGEMM, TF32 off) and on the CPU (what the reference computes with)
against the fp64 product.  Reports the rms and worst error relative to rms(D), and the error's correlation with
-sign(D): an accumulator that TRUNCATES (round toward zero) instead of rounding to nearest shows up as a bias
towards zero that grows linearly with the number of accumulation steps.

    python scripts/gemm_precision.py > gpurun_out/gemm_precision.json
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lbic_b200  # noqa: E402
from lbic_b200 import weights  # noqa: E402
from lbic_b200.net import BlockBasedImgCompLossyNetv9  # noqa: E402


def stats(D, ref):
    e = D.double() - ref
    rms = ref.pow(2).mean().sqrt()
    return dict(rms_rel=float(e.pow(2).mean().sqrt() / rms), max_rel=float(e.abs().max() / rms),
                bias_toward_zero=float((-(e * ref.sign())).mean() / rms))


def main():
    dev = torch.device("cuda", 0)
    torch.backends.cuda.matmul.allow_tf32 = False
    cfg = lbic_b200.load_config("B8_lowrate")
    m = BlockBasedImgCompLossyNetv9(cfg, device=dev)
    m.load_state_dict(weights.synth_state_dict(cfg, 1337))
    m.update(force=True)
    out = []
    for K, C, kind in [(576, 96, "randn"), (768, 768, "randn"), (960, 768, "randn"), (3840, 768, "randn"),
                       (576, 96, "positive"), (768, 768, "positive")]:
        g = torch.Generator().manual_seed(K * 7 + C)
        A = torch.randn(512, K, generator=g)
        W = torch.randn(C, K, generator=g) / K ** 0.5
        if kind == "positive":          # GDN-like: squares times non-negative gamma (no cancellation, large accumulator)
            A, W = A * A, W.abs()
        # the product path stores each layer's weights times 2^k so that max |w| lies in [2^9, 2^10) (api.cu finish_layer)
        k = 9 - int(torch.floor(torch.log2(W.abs().max())))
        W = W * 2.0 ** k
        ref = A.double() @ W.double().T
        rec = dict(K=K, cout=C, operands=kind)
        Ad, Wd = A.to(dev), W.to(dev)
        for core in ("tcgen05", "simt"):
            m.set_gemm_core(core)
            rec[core] = stats(m.debug_gemm(Ad, Wd).cpu(), ref)
        m.set_gemm_core("tcgen05")
        rec["cublas_f32"] = stats((Ad @ Wd.T).cpu(), ref)
        rec["cpu_f32"] = stats(A @ W.T, ref)
        # the hi/lo representation error alone: fp64 product of the split operands (hi + lo of each)
        def split(t):
            hi = t.half().float()
            return (hi + (t - hi).half().float()).double()
        rec["hilo_exact"] = stats(split(A) @ split(W).T, ref)
        out.append(rec)
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
