"""Is an isolated layer launch DRAM-bound?  Same layer (K = 768, N = 512, CTA pairs, 256-wide tiles) on a problem whose
operands + outputs fit the 126 MB L2 (one wave of 74 tiles, buffers reused by 30 back-to-back launches) and on one 4x
as large (four waves, 270 MB per launch): time per wave."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lbic_b200
from lbic_b200 import _lib
from lbic_b200.net import BlockBasedImgCompLossyNetv9
from lbic_b200.weights import synth_state_dict
cfg = lbic_b200.load_config("B8_lowrate")
m = BlockBasedImgCompLossyNetv9(cfg, device="cuda:0")
m.load_state_dict(synth_state_dict(cfg)); m.update()
L = _lib.lib()
def opt(o, v): _lib.check(L.lbic_set_option(m._need(), o, v))
opt(_lib.LBIC_OPT_WS, 2); opt(_lib.LBIC_OPT_PAIR, 3)
names = ["RAW", "PREGDN", "GDN", "QUANT", "LRELU", "KSI", "RECON"]
for epi in (0, 4, 1, 2):
    for waves in (1, 2, 4, 8):
        R = 74 * 256 * waves // 2
        ms = ctypes.c_double()
        _lib.check(L.lbic_debug_gemm_bench(m._need(), R, 768, 512, epi, 30, ctypes.byref(ms)))
        mb = R * (768 * 4 + 512 * (8 if epi == 1 else 4) + (512 * 4 if epi == 2 else 0)) / 1e6
        print(f"{names[epi]:7s} R={R:6d} ({waves} waves of 74 tiles, {mb:6.0f} MB touched per launch): {ms.value*1e3:7.1f} us = {ms.value*1e3/waves:6.1f} us per wave", flush=True)
