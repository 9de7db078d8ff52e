"""Isolated GEMM + epilogue timing per epilogue mode (persistent kernel, pair form)."""
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
def opt(o, v): _lib.check(L.lbic_set_option(m._need(), o, v))
names = ["RAW", "PREGDN", "GDN", "QUANT", "LRELU", "KSI", "RECON"]
def run(R, K, C, epi, pair, ws=2, iters=30):
    opt(_lib.LBIC_OPT_WS, ws); opt(_lib.LBIC_OPT_PAIR, pair)
    ms = ctypes.c_double()
    _lib.check(L.lbic_debug_gemm_bench(m._need(), R, K, C, epi, iters, ctypes.byref(ms)))
    tf = 2.0 * R * K * C / (ms.value * 1e-3) / 1e12
    print(f"ws={ws} pair={pair} R={R:6d} K={K:5d} C={C:4d} {names[epi]:7s}: {ms.value*1e3:8.1f} us  {tf:7.1f} TF/s alg ({3*tf:7.1f} mma)", flush=True)
R = int(sys.argv[1]) if len(sys.argv) > 1 else 24576
for pair in (1, 3):
    for epi in range(7):
        run(R, 768, 768, epi, pair)
    for epi in (1, 2):
        run(R, 672, 672, epi, pair)
    run(R, 576, 96, 3, pair); run(R, 576, 96, 1, pair); run(R, 576, 96, 0, pair)
    run(R, 576, 192, 6, pair); run(R, 768, 192, 5, pair)
    run(R, 768, 1152, 4, pair); run(R, 1152, 960, 4, pair); run(R, 960, 768, 4, pair); run(R, 960, 768, 1, pair); run(R, 768, 672, 1, pair); run(R, 672, 576, 1, pair); run(R, 576, 576, 2, pair)
run(R, 576, 96, 3, 0, ws=0); run(R, 576, 192, 6, 0, ws=0); run(R, 768, 192, 5, 0, ws=0)
