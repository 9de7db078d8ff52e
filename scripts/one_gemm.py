"""One isolated GEMM + epilogue configuration (for ncu): one_gemm.py R K C epi pair [iters]"""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lbic_b200
from lbic_b200 import _lib
from lbic_b200.net import BlockBasedImgCompLossyNetv9
from lbic_b200.weights import synth_state_dict
cfg = lbic_b200.load_config("B8_lowrate")
m = BlockBasedImgCompLossyNetv9(cfg, device="cuda:0")
m.load_state_dict(synth_state_dict(cfg)); m.update()
L = _lib.lib()
R, K, C, epi, pair = [int(a) for a in sys.argv[1:6]]
iters = int(sys.argv[6]) if len(sys.argv) > 6 else 3
_lib.check(L.lbic_set_option(m._need(), _lib.LBIC_OPT_WS, 2))
_lib.check(L.lbic_set_option(m._need(), _lib.LBIC_OPT_PAIR, pair))
ms = ctypes.c_double()
_lib.check(L.lbic_debug_gemm_bench(m._need(), R, K, C, epi, iters, ctypes.byref(ms)))
print(f"R={R} K={K} C={C} epi={epi} pair={pair}: {ms.value*1e3:.1f} us")
