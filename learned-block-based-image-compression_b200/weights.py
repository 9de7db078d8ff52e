"""state_dict layout of the reference's v9 network and deterministic synthetic weights.

The key set / shapes follow the reference model constructor
(graphs/models/BlockBasedImgCompLossy_net.py:259-317; MaskedConv2d registers a `mask` buffer,
graphs/layers/masked_conv2d.py:5-17; GDN holds beta/gamma plus four reparam buffers,
graphs/layers/gdn_compressai.py:43-63, utils/parametrizers.py:33-40).  Release checkpoints are
not available offline, so tests and bench.py use `synth_state_dict`: "random-init weights of the
named architecture", drawn per key from its own seeded generator (independent of module
construction order) and then conditioned so the closed loop is actually exercised
(raw default init makes every symbol 0).
"""
from __future__ import annotations

import math
import zlib
from collections import OrderedDict

import torch

REPARAM_OFFSET = 2.0 ** -18
PEDESTAL = REPARAM_OFFSET ** 2
BETA_MIN = 1e-6


def widths(cfg):
    """Channel widths of the three sub-networks (NET:262-302)."""
    B, N, M = int(cfg.block_size), int(cfg.N), int(cfg.M)
    cin = 3 * B * B
    return dict(cin=cin, N=N, C2=N // 8 * 7, C3=N // 8 * 6, M=M,
                E1=N // 8 * 12, E2=N // 8 * 10, E3=N // 8 * 8, EO=2 * M,
                k0=int(cfg.KS[0]), k1=int(cfg.KS[1]))


def conv_mask(mask_type: str, cout: int, cin: int, k: int) -> torch.Tensor:
    """MaskedConv2d mask (masked_conv2d.py:9-17)."""
    m = torch.ones(cout, cin, k, k)
    if k > 1:
        m[:, :, k // 2, k // 2 + (mask_type == "B"):] = 0
        m[:, :, k // 2 + 1:] = 0
    elif mask_type == "A":
        m[:, :, 0, 0:] = 0
    return m


def layer_specs(cfg):
    """Ordered list of (prefix, kind, params) describing every module with state."""
    w = widths(cfg)
    cin, N, C2, C3, M = w["cin"], w["N"], w["C2"], w["C3"], w["M"]
    E1, E2, E3, EO, k0, k1 = w["E1"], w["E2"], w["E3"], w["EO"], w["k0"], w["k1"]
    L = []
    L.append(("prtr_forward1", "conv", ("B", N, cin, 1)))
    L.append(("prtr_forward2", "conv", ("A", N, cin, k0)))
    L.append(("prtr_forward3.0", "gdn", (N, False)))
    L.append(("prtr_forward3.1", "conv", ("B", C2, N, 1)))
    L.append(("prtr_forward3.2", "gdn", (C2, False)))
    L.append(("prtr_forward3.3", "conv", ("B", C3, C2, 1)))
    L.append(("prtr_forward3.4", "gdn", (C3, False)))
    L.append(("prtr_forward3.5", "conv", ("B", M, C3, 1)))
    L.append(("prtr_inverse1", "conv", ("B", N, M, 1)))
    L.append(("prtr_inverse2", "conv", ("A", N, cin, k0)))
    L.append(("prtr_inverse3.0", "gdn", (N, True)))
    L.append(("prtr_inverse3.1", "conv", ("B", C2, N, 1)))
    L.append(("prtr_inverse3.2", "gdn", (C2, True)))
    L.append(("prtr_inverse3.3", "conv", ("B", C3, C2, 1)))
    L.append(("prtr_inverse3.4", "gdn", (C3, True)))
    L.append(("prtr_inverse3.5", "conv", ("B", cin, C3, 1)))
    L.append(("get_meanscale.0", "conv", ("A", E1, cin, k0)))
    L.append(("get_meanscale.2", "conv", ("B", E2, E1, k1)))
    L.append(("get_meanscale.4", "conv", ("B", E3, E2, 1)))
    L.append(("get_meanscale.6", "conv", ("B", EO, E3, 1)))
    return L


def _gen(seed: int, key: str) -> torch.Generator:
    g = torch.Generator()
    g.manual_seed((int(seed) * 1000003 + zlib.crc32(key.encode())) % (2 ** 63))
    return g


def reparam_init(x: torch.Tensor) -> torch.Tensor:
    """NonNegativeParametrizer.init (utils/parametrizers.py:42-43)."""
    ped = torch.tensor([PEDESTAL], dtype=torch.float32)
    return torch.sqrt(torch.max(x + ped, ped))


def synth_state_dict(cfg, seed: int = 1337, conditioned: bool = True, harsh: bool = False, latent_gain: float = 60.0,
                     scale_span: float = 5.2):
    """Synthetic state_dict with the reference's exact key set, order, shapes and dtypes.

    conv weight/bias ~ U(-1/sqrt(fan_in), 1/sqrt(fan_in)) (PyTorch's default Conv2d scale); masked
    taps are left holding random garbage on purpose (the reference multiplies by `mask` at use,
    NET:381).  conditioned=True applies the SURVEY.md section 8(d) recipe plus a dense non-negative
    GDN gamma so that the six gamma-matmuls are exercised (default-init gamma is exactly diagonal).
    harsh=True keeps prtr_inverse1 at full scale (clamp active on about half the samples).
    latent_gain / scale_span set the rate of the synthetic model: the last encoder layer is scaled by latent_gain and the
    predicted scales are spread over exp(U(0, scale_span) - 2.2); the defaults (60, 5.2) give ~11 bpp on noise-like
    images (every golden fixture uses them), (2.5, 2.0) about 1 bpp, the range of the reference's released models.
    """
    w = widths(cfg)
    M = w["M"]
    sd = OrderedDict()
    for prefix, kind, p in layer_specs(cfg):
        if kind == "conv":
            mtype, cout, cin, k = p
            bound = 1.0 / math.sqrt(cin * k * k)
            g = _gen(seed, prefix)
            wt = (torch.rand(cout, cin, k, k, generator=g) * 2 - 1) * bound
            bs = (torch.rand(cout, generator=g) * 2 - 1) * bound
            sd[prefix + ".weight"] = wt
            sd[prefix + ".bias"] = bs
            sd[prefix + ".mask"] = conv_mask(mtype, cout, cin, k)
        else:
            C, _inv = p
            g = _gen(seed, prefix)
            if conditioned:
                beta = 0.5 + torch.rand(C, generator=g)
                gamma = 0.1 * torch.eye(C) + (0.6 / C) * torch.rand(C, C, generator=g)
            else:
                beta = torch.ones(C)
                gamma = 0.1 * torch.eye(C)
            sd[prefix + ".beta"] = reparam_init(beta)
            sd[prefix + ".gamma"] = reparam_init(gamma)
            sd[prefix + ".beta_reparam.pedestal"] = torch.tensor([PEDESTAL], dtype=torch.float32)
            sd[prefix + ".beta_reparam.lower_bound.bound"] = torch.tensor(
                [(BETA_MIN + PEDESTAL) ** 0.5], dtype=torch.float32)
            sd[prefix + ".gamma_reparam.pedestal"] = torch.tensor([PEDESTAL], dtype=torch.float32)
            sd[prefix + ".gamma_reparam.lower_bound.bound"] = torch.tensor(
                [(0.0 + PEDESTAL) ** 0.5], dtype=torch.float32)
    if conditioned:
        sd["prtr_forward3.5.weight"] *= latent_gain
        sd["prtr_forward3.5.bias"] *= latent_gain
        if not harsh:
            sd["prtr_inverse1.weight"] *= 0.3
        sd["get_meanscale.6.weight"] *= 20.0
        g = _gen(seed, "scale_bias")
        sd["get_meanscale.6.bias"][:M] = torch.exp(torch.rand(M, generator=g) * scale_span - 2.2)
    # entropy-model buffers: empty until update(), exactly as in released checkpoints (AGENT:570)
    sd["conditional_gaussian_model._offset"] = torch.IntTensor()
    sd["conditional_gaussian_model._quantized_cdf"] = torch.IntTensor()
    sd["conditional_gaussian_model._cdf_length"] = torch.IntTensor()
    sd["conditional_gaussian_model.scale_table"] = torch.Tensor()
    sd["conditional_gaussian_model.scale_bound"] = torch.tensor([0.11])
    sd["conditional_gaussian_model.likelihood_lower_bound.bound"] = torch.tensor([1e-9])
    sd["conditional_gaussian_model.lower_bound_scale.bound"] = torch.tensor([0.11])
    return sd


def synth_postpm_state_dict(cfg, seed: int = 4321, gain: float = 8.0):
    """Deterministic weights for the optional post-processing module (BlkBasedPostProcessing, NET:455-476), with the
    reference module's key set: res_net.0 = Conv2d(C, 4C, 3), res_net.2 = Conv2d(4C, C, 1), C = 3B^2.  PyTorch's default
    Conv2d scale, the last layer times `gain` so that the residual is visible (default init gives ~1e-3)."""
    C = 3 * int(cfg.block_size) ** 2
    sd = OrderedDict()
    for name, cout, cin, k, g_ in (("res_net.0", 4 * C, C, 3, 1.0), ("res_net.2", C, 4 * C, 1, gain)):
        bound = 1.0 / math.sqrt(cin * k * k)
        g = _gen(seed, "postpm." + name)
        sd[name + ".weight"] = (torch.rand(cout, cin, k, k, generator=g) * 2 - 1) * bound * g_
        sd[name + ".bias"] = (torch.rand(cout, generator=g) * 2 - 1) * bound
    return sd


def synth_images(n: int, H: int, W: int, seed0: int = 1000, kind: str = "smooth") -> torch.Tensor:
    """Synthetic test images in [0,1], (n,3,H,W) fp32: low-resolution
    noise bicubic-upsampled plus 0.05*randn ("smooth"), or white noise ("noise").  Image i uses
    seed seed0+i."""
    out = torch.empty(n, 3, H, W)
    for i in range(n):
        g = torch.Generator()
        g.manual_seed(seed0 + i)
        if kind == "noise":
            out[i] = torch.rand(3, H, W, generator=g)
        else:
            low = torch.rand(1, 3, max(H // 16, 2), max(W // 16, 2), generator=g)
            up = torch.nn.functional.interpolate(low, size=(H, W), mode="bicubic", align_corners=False)
            out[i] = (up[0] + 0.05 * torch.randn(3, H, W, generator=g)).clamp_(0, 1)
    return out


def synth_image_u8(H: int, W: int, seed: int = 0):
    """Bit-reproducible synthetic 8-bit RGB image (numpy integer arithmetic only: identical on every machine):
    a coarse random grid, bilinearly interpolated in fixed point, plus +-12 levels of hash noise.
    Returns a uint8 array (3, H, W)."""
    import numpy as np
    rng = np.random.RandomState(seed)        # MT19937 integer draws are platform independent
    gh, gw = H // 16 + 2, W // 16 + 2
    grid = rng.randint(0, 256, size=(3, gh, gw)).astype(np.int64)
    ys, xs = np.arange(H, dtype=np.int64), np.arange(W, dtype=np.int64)
    y0, fy = ys // 16, ys % 16
    x0, fx = xs // 16, xs % 16
    g00 = grid[:, y0][:, :, x0]
    g01 = grid[:, y0][:, :, x0 + 1]
    g10 = grid[:, y0 + 1][:, :, x0]
    g11 = grid[:, y0 + 1][:, :, x0 + 1]
    wy, wx = fy[None, :, None], fx[None, None, :]
    smooth = ((16 - wy) * ((16 - wx) * g00 + wx * g01) + wy * ((16 - wx) * g10 + wx * g11)) // 256
    noise = rng.randint(-12, 13, size=(3, H, W)).astype(np.int64)
    return np.clip(smooth + noise, 0, 255).astype(np.uint8)


def u8_to_model_input(img_u8):
    """uint8 (3,H,W) or (n,3,H,W) -> float32 tensor in [-0.5, 0.5] exactly as eval_model feeds it (x/255 - 0.5)."""
    import numpy as np
    a = np.asarray(img_u8)
    if a.ndim == 3:
        a = a[None]
    return torch.from_numpy(a.astype(np.float32) / np.float32(255.0) - np.float32(0.5))
