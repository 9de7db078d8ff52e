"""Golden for the optional post-processing module: the UNMODIFIED reference class BlkBasedPostProcessing
(graphs/models/BlockBasedImgCompLossy_net.py:455-476, through oracle/ref_shim) with the deterministic weights of
lbic_b200.weights.synth_postpm_state_dict loaded into it, applied to the closed-loop reconstructions of the small golden
cases (only the outputs are stored; the test regenerates the weights).  Asserts oracle.nets.postprocess == reference first.
    python tests/golden/make_golden_postpm.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import lbic_b200  # noqa: E402
from lbic_b200 import weights  # noqa: E402
from oracle import nets  # noqa: E402
from oracle.ref_shim import load_reference  # noqa: E402


def main():
    ref = load_reference.load()
    for case in ("B8_lowrate_6x9", "B4_highrate_7x10", "B16_lowrate_3x5"):
        c = np.load(os.path.join(HERE, f"case_{case}.npz"))
        cfg = lbic_b200.load_config(str(c["config"]))
        sd = weights.synth_postpm_state_dict(cfg, 4321)
        pm = ref.BlkBasedPostProcessing(cfg).eval()
        pm.load_state_dict(sd)                             # strict: the key set is the reference module's
        z = torch.from_numpy(c["zhat"])
        with torch.no_grad():
            out = pm(z)
        mine = nets.postprocess(sd, z)
        assert torch.equal(mine, out), "oracle restatement differs from the reference module"
        np.savez_compressed(os.path.join(HERE, f"postpm_{case}.npz"), config=str(c["config"]), seed=4321, out=out.numpy())
        print(case, tuple(out.shape), "residual max", float((out - z).abs().max()))


if __name__ == "__main__":
    main()
