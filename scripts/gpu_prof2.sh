#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --images 64 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/prof2_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_chain -s 250 -c 2 -o gpurun_out/prof_chain_r1 $CMD > gpurun_out/ncu_chain.log 2>&1
tail -3 gpurun_out/ncu_chain.log
python - <<'PY' > gpurun_out/cluster_sweep.log 2>&1
import sys, os, time, torch
sys.path.insert(0, os.getcwd())
import lbic_b200
from lbic_b200 import weights
from lbic_b200.layout import arrange_block_pixels_to_channel_dim
from lbic_b200.net import BlockBasedImgCompLossyNetv9
cfg = lbic_b200.load_config("B8_lowrate")
dev = torch.device("cuda:0")
m = BlockBasedImgCompLossyNetv9(cfg, device=dev)
m.load_state_dict(weights.synth_state_dict(cfg, 1337)); m.update(force=True)
for n in (64, 256):
    img = torch.rand(n, 3, 512, 768, device=dev)
    x = arrange_block_pixels_to_channel_dim(img - 0.5, 8)
    out = m.encode_device(x, lanes=0)
    for chain, S in ((0, 0), (1, 0), (1, 1), (1, 2), (1, 3), (1, 4), (1, 6), (1, 8)):
        m.set_option("chain", chain); m.set_option("cluster", S)
        m.encode_device(x, lanes=0, out=out); torch.cuda.synchronize()
        t = time.perf_counter()
        m.encode_device(x, lanes=0, out=out); torch.cuda.synchronize()
        dt = time.perf_counter() - t
        print(f"n={n} chain={chain} S={S}: encode {dt*1e3:.1f} ms  {n*512*768/dt/1e6:.1f} Mpix/s", flush=True)
PY
cat gpurun_out/cluster_sweep.log
