"""Tile-kernel timing sweep on the GPU box (tuning aid; prints ms per launch and TFLOP/s algorithmic)."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lbic_b200
from lbic_b200 import _lib
from lbic_b200.net import BlockBasedImgCompLossyNetv9
m = BlockBasedImgCompLossyNetv9(lbic_b200.load_config("B8_lowrate"), device="cuda:0")
L = _lib.lib()
def run(R, K, C, epi, bn=0, iters=50):
    _lib.check(L.lbic_set_option(m._need(), _lib.LBIC_OPT_FORCE_BN, bn))
    ms = ctypes.c_double()
    _lib.check(L.lbic_debug_gemm_bench(m._need(), R, K, C, epi, iters, ctypes.byref(ms)))
    tf = 2.0 * R * K * C / (ms.value * 1e-3) / 1e12
    print(f"R={R:6d} K={K:5d} C={C:4d} epi={epi} bn={bn:3d}: {ms.value*1e3:8.1f} us  {tf:7.1f} TF/s alg ({3*tf:7.1f} tensor)", flush=True)
for epi in (0, 1):
    for K in (64, 256, 768, 1536, 3072):
        run(3072, K, 768, epi)
for bn in (64, 128, 192, 256):
    for K in (256, 768, 3072):
        run(3072, K, 768, 1, bn)
for R in (128, 1024, 6144, 18944, 18944 * 2, 18944 * 4):
    run(R, 768, 768, 1)
    run(R, 768, 768, 1, 128)
