"""CTA-pair kernel: bit-identity against the per-tile kernel on raw GEMMs, then a timing sweep."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lbic_b200
from lbic_b200 import _lib
from lbic_b200.net import BlockBasedImgCompLossyNetv9
m = BlockBasedImgCompLossyNetv9(lbic_b200.load_config("B8_lowrate"), device="cuda:0")
L = _lib.lib()
def opt(o, v): _lib.check(L.lbic_set_option(m._need(), o, v))
torch.manual_seed(0)
bad = 0
for (R, K, C) in [(256, 64, 192), (512, 192, 96), (1000, 768, 672), (4096 + 130, 960, 768), (300, 576, 192), (20000, 1152, 960)]:
    A = torch.randn(R, K, device="cuda"); W = torch.randn(C, K, device="cuda") / K ** 0.5
    opt(_lib.LBIC_OPT_WS, 0); opt(_lib.LBIC_OPT_PAIR, 0)
    d0 = m.debug_gemm(A, W)
    opt(_lib.LBIC_OPT_WS, 2)
    d1 = m.debug_gemm(A, W)
    opt(_lib.LBIC_OPT_PAIR, 1)
    d2 = m.debug_gemm(A, W)
    e1 = (d1 != d0).sum().item(); e2 = (d2 != d0).sum().item()
    print(f"R={R} K={K} C={C}: ws mismatches {e1}, pair mismatches {e2}, max|pair-tc| {(d2-d0).abs().max().item():.3e}", flush=True)
    bad += e1 + e2
print("BIT-IDENTICAL" if bad == 0 else "MISMATCH", flush=True)
if bad == 0 and len(sys.argv) > 1:
    def run(R, K, C, epi, pair, iters=30):
        opt(_lib.LBIC_OPT_WS, 2); opt(_lib.LBIC_OPT_PAIR, pair)
        ms = ctypes.c_double()
        _lib.check(L.lbic_debug_gemm_bench(m._need(), R, K, C, epi, iters, ctypes.byref(ms)))
        tf = 2.0 * R * K * C / (ms.value * 1e-3) / 1e12
        print(f"pair={pair} R={R:6d} K={K:5d} C={C:4d} epi={epi}: {ms.value*1e3:8.1f} us  {tf:7.1f} TF/s alg ({3*tf:7.1f} mma)", flush=True)
    for pair in (0, 1):
        for (K, C) in [(768, 1152), (1152, 960), (960, 768), (768, 768), (768, 672), (672, 672), (672, 576), (576, 576), (576, 96), (576, 192), (768, 192)]:
            run(24576, K, C, 1, pair)
        run(37888, 1536, 768, 1, pair)
        run(37888, 3072, 768, 0, pair)
