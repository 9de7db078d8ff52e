"""ws-kernel vs per-tile kernel timing sweep (tuning aid)."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lbic_b200
from lbic_b200 import _lib
from lbic_b200.net import BlockBasedImgCompLossyNetv9
m = BlockBasedImgCompLossyNetv9(lbic_b200.load_config("B8_lowrate"), device="cuda:0")
L = _lib.lib()
def run(R, K, C, epi, ws, bn=0, iters=30):
    _lib.check(L.lbic_set_option(m._need(), _lib.LBIC_OPT_FORCE_BN, bn))
    _lib.check(L.lbic_set_option(m._need(), _lib.LBIC_OPT_WS, 2 if ws else 0))
    ms = ctypes.c_double()
    _lib.check(L.lbic_debug_gemm_bench(m._need(), R, K, C, epi, iters, ctypes.byref(ms)))
    tf = 2.0 * R * K * C / (ms.value * 1e-3) / 1e12
    print(f"ws={ws} R={R:6d} K={K:5d} C={C:4d} epi={epi} bn={bn:3d}: {ms.value*1e3:8.1f} us  {tf:7.1f} TF/s alg ({3*tf:7.1f} tensor)", flush=True)
for ws in (0, 1):
    for K in (256, 768, 1536, 3072):
        run(37888, K, 768, 1, ws)
    for bn in (128, 160, 192):
        run(37888, 768, 768, 1, ws, bn)
    run(37888, 768, 768, 0, ws)
    run(75776, 768, 768, 1, ws)
    run(75776, 960, 768, 1, ws)
    run(75776, 768, 576, 1, ws)
