"""CPU restatement (torch fp32, CPU) of the reference's v9 closed-loop codec -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product path (learned-block-based-image-compression_b200) never does.

Every function cites the reference lines it follows (paths relative to /root/reference):
  NET   = graphs/models/BlockBasedImgCompLossy_net.py
  MC    = graphs/layers/masked_conv2d.py
  GDNF  = graphs/layers/gdn_compressai.py
  ENT   = graphs/layers/entropy_layers_cai.py
  AGENT = agents/blkbsdimgcomp_agent.py
This restatement is pinned against the reference's own code executed through oracle/ref_shim
(tests/golden/make_golden.py, tests/test_oracle_vs_reference.py); the reference itself ships no
golden vectors.
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn.functional as F

from . import native

SCALES_MIN, SCALES_MAX, SCALES_LEVELS = 0.11, 256, 64   # NET:13-15


def get_scale_table():
    """NET:17-18."""
    return torch.exp(torch.linspace(math.log(SCALES_MIN), math.log(SCALES_MAX), SCALES_LEVELS))


# ----------------------------------------------------------------------------------------------
# entropy-model tables  (ENT:590-613 GaussianConditional.update, ENT:175-183 _pmf_to_cdf)
# ----------------------------------------------------------------------------------------------
def build_tables(scale_table=None, tail_mass: float = 1e-9, precision: int = 16):
    """Returns (quantized_cdf [64,3133] int32, cdf_length [64] int32, offset [64] int32)."""
    import scipy.stats

    st = get_scale_table() if scale_table is None else torch.as_tensor(scale_table, dtype=torch.float32)
    multiplier = -scipy.stats.norm.ppf(tail_mass / 2)                    # ENT:575-577, 591
    pmf_center = torch.ceil(st * multiplier).int()                       # ENT:592
    pmf_length = 2 * pmf_center + 1
    max_length = int(torch.max(pmf_length).item())
    samples = torch.abs(torch.arange(max_length).int() - pmf_center[:, None]).float()
    scale = st.unsqueeze(1).float()

    def cumulative(v):                                                   # ENT:569-573
        return 0.5 * torch.erfc(float(-(2 ** -0.5)) * v)

    upper = cumulative((0.5 - samples) / scale)
    lower = cumulative((-0.5 - samples) / scale)
    pmf = upper - lower
    tail = 2 * lower[:, :1]
    cdf = torch.zeros((len(pmf_length), max_length + 2), dtype=torch.int32)
    for i in range(len(pmf_length)):                                     # ENT:175-183
        prob = torch.cat((pmf[i, : int(pmf_length[i])], tail[i]), dim=0)
        c = native.pmf_to_quantized_cdf(prob.tolist(), precision)
        cdf[i, : len(c)] = torch.from_numpy(c.astype(np.int32))
    return cdf, (pmf_length + 2).int(), (-pmf_center).int()


# ----------------------------------------------------------------------------------------------
# effective parameters
# ----------------------------------------------------------------------------------------------
def _nonneg(p, bound, pedestal):
    """NonNegativeParametrizer.forward (utils/parametrizers.py:45-48)."""
    return torch.max(p, bound) ** 2 - pedestal


def effective_params(sd, cfg):
    """Masked conv weights (NET:381 `weight * mask`) and re-parametrised GDN beta/gamma (GDNF:68-69)."""
    P = SimpleNamespace()
    sd = {k: v.detach().float().cpu() if v.dtype.is_floating_point else v.detach().cpu() for k, v in sd.items()}

    def conv(prefix):
        return (sd[prefix + ".weight"] * sd[prefix + ".mask"], sd[prefix + ".bias"])

    def gdn(prefix):
        beta = _nonneg(sd[prefix + ".beta"], sd[prefix + ".beta_reparam.lower_bound.bound"],
                       sd[prefix + ".beta_reparam.pedestal"])
        gamma = _nonneg(sd[prefix + ".gamma"], sd[prefix + ".gamma_reparam.lower_bound.bound"],
                        sd[prefix + ".gamma_reparam.pedestal"])
        return (beta, gamma)

    P.f1, P.f2 = conv("prtr_forward1"), conv("prtr_forward2")
    P.f3 = [gdn("prtr_forward3.0"), conv("prtr_forward3.1"), gdn("prtr_forward3.2"),
            conv("prtr_forward3.3"), gdn("prtr_forward3.4"), conv("prtr_forward3.5")]
    P.i1, P.i2 = conv("prtr_inverse1"), conv("prtr_inverse2")
    P.i3 = [gdn("prtr_inverse3.0"), conv("prtr_inverse3.1"), gdn("prtr_inverse3.2"),
            conv("prtr_inverse3.3"), gdn("prtr_inverse3.4"), conv("prtr_inverse3.5")]
    P.e = [conv("get_meanscale.0"), conv("get_meanscale.2"), conv("get_meanscale.4"), conv("get_meanscale.6")]
    P.scale_table = get_scale_table()
    P.M = int(cfg.M)
    P.LRU = sum(int(k) // 2 for k in cfg.KS)                             # AGENT:481-489 'compress'
    return P


def _gdn(x, beta_gamma, inverse):
    """GDNF:65-80."""
    beta, gamma = beta_gamma
    C = x.shape[1]
    norm = F.conv2d(x ** 2, gamma.reshape(C, C, 1, 1), beta)
    norm = torch.sqrt(norm) if inverse else torch.rsqrt(norm)
    return x * norm


def _chain(x, layers, inverse):
    """nn.Sequential(GDN, conv1x1, GDN, conv1x1, GDN, conv1x1)  NET:275-282 / NET:286-293."""
    for i, p in enumerate(layers):
        x = _gdn(x, p, inverse) if i % 2 == 0 else F.conv2d(x, p[0], p[1])
    return x


def forward_prtr(P, zhat_win3, x_blk):
    """NET:379-382 (valid conv over the 3x3 window + 1x1 on x)."""
    return _chain(F.conv2d(x_blk, *P.f1) + F.conv2d(zhat_win3, *P.f2), P.f3, False)


def inverse_prtr(P, zhat_win3, y_qnt):
    """NET:384-387."""
    return _chain(F.conv2d(y_qnt, *P.i1) + F.conv2d(zhat_win3, *P.i2), P.i3, True)


def get_meanscale(P, zhat_win):
    """NET:389-398: four valid (padding=0) masked convs with LeakyReLU(0.01) between."""
    z = F.leaky_relu(F.conv2d(zhat_win, *P.e[0]))
    z = F.leaky_relu(F.conv2d(z, *P.e[1]))
    z = F.leaky_relu(F.conv2d(z, *P.e[2]))
    return F.conv2d(z, *P.e[3])


def build_indexes(P, scales):
    """ENT:649-654 with LowerBound(0.11) (ENT:553, utils/bound_ops.py:22-23)."""
    s = torch.max(scales, torch.tensor([SCALES_MIN]))
    idx = torch.full(s.shape, SCALES_LEVELS - 1, dtype=torch.int32)
    for t in P.scale_table[:-1]:
        idx -= (s <= t).int()
    return idx


def quantize_symbols(y, means):
    """ENT:126-151 mode "symbols": round-half-even of (y - means) -> int32."""
    return torch.round(y - means).int()


# ----------------------------------------------------------------------------------------------
# the closed loop, block by block in raster order
# ----------------------------------------------------------------------------------------------
def _window(zhat, v, h, r):
    """Zero-padded (2r+1)x(2r+1) window of zhat centred on block (v,h)   NET:340-351."""
    _, _, hg, wd = zhat.shape
    UU, BB = max(0, v - r), min(hg, v + r + 1)
    LL, RR = max(0, h - r), min(wd, h + r + 1)
    pads = (r - h + LL, h + r + 1 - RR, r - v + UU, v + r + 1 - BB)
    win = zhat[:, :, UU:BB, LL:RR]
    if any(p > 0 for p in pads):
        win = F.pad(win, pads, mode="constant", value=0.0)
    return win


@torch.no_grad()
def compress_loop(P, x, max_blocks=None):
    """NET:319-357 + NET:363-377 without the entropy coder.
    x: (1, 3B^2, Hb, Wb) in [-0.5, 0.5].  Returns symbols (Hb,Wb,M) int32, indexes (Hb,Wb,M) int32,
    zhat (1,3B^2,Hb,Wb).  max_blocks bounds the work (bench sampling)."""
    _, _, hg, wd = x.shape
    r = P.LRU
    zhat = torch.zeros_like(x)
    syms = torch.zeros(hg, wd, P.M, dtype=torch.int32)
    idxs = torch.zeros(hg, wd, P.M, dtype=torch.int32)
    done = 0
    for v in range(hg):
        for h in range(wd):
            win = _window(zhat, v, h, r)
            c = win.shape[3] // 2
            win3 = win[:, :, c - 1:c + 2, c - 1:c + 2]
            y = forward_prtr(P, win3, x[:, :, v:v + 1, h:h + 1])
            ksi = get_meanscale(P, win)
            scales, means = ksi.chunk(2, dim=1)                          # NET:369 (scales first)
            idx = build_indexes(P, scales)
            sym = quantize_symbols(y, means)
            y_qnt = sym + means                                          # NET:374
            xhat = inverse_prtr(P, win3, y_qnt)
            syms[v, h] = sym[0, :, 0, 0]
            idxs[v, h] = idx[0, :, 0, 0]
            zhat[:, :, v, h] = xhat[:, :, 0, 0].clamp_(-0.5, 0.5)        # NET:357
            done += 1
            if max_blocks is not None and done >= max_blocks:
                return syms, idxs, zhat
    return syms, idxs, zhat


@torch.no_grad()
def decompress_loop(P, tables, bitstream, xshape, max_blocks=None):
    """NET:400-452: per block entropy-net -> indexes -> decode M symbols -> dequantise -> decoder net."""
    _, ch, hg, wd = xshape
    r = P.LRU
    T = native.Tables(*[t.numpy() for t in tables])
    dec = native.RansDecoder()
    dec.set_stream(bitstream)
    zhat = torch.zeros(1, ch, hg, wd)
    done = 0
    for v in range(hg):
        for h in range(wd):
            win = _window(zhat, v, h, r)
            ksi = get_meanscale(P, win)
            scales, means = ksi.chunk(2, dim=1)
            idx = build_indexes(P, scales[:, :, 0, 0])
            rv = dec.decode_stream(idx.reshape(-1).numpy(), T)
            rv = torch.from_numpy(rv.astype(np.float32)).reshape(1, -1)   # NET:440
            y_qnt = (rv + means[:, :, 0, 0]).reshape(1, -1, 1, 1)         # ENT:159-168
            c = win.shape[3] // 2
            xhat = inverse_prtr(P, win[:, :, c - 1:c + 2, c - 1:c + 2], y_qnt)
            zhat[:, :, v, h] = xhat[:, :, 0, 0].clamp_(-0.5, 0.5)         # NET:450
            done += 1
            if max_blocks is not None and done >= max_blocks:
                return zhat
    return zhat


def compress(P, tables, x):
    """Full NET:319-361: loop + rANS over all symbols (raster block order, channels inner NET:353-354)."""
    syms, idxs, zhat = compress_loop(P, x)
    T = native.Tables(*[t.numpy() for t in tables])
    stream = native.rans_encode(syms.reshape(-1).numpy(), idxs.reshape(-1).numpy(), T)
    return stream, zhat, syms, idxs


# ----------------------------------------------------------------------------------------------
# whole-image (open-loop) restatement -- SURVEY.md Appendix A.6: the closed loop's result is the
# unique fixed point of one fully parallel evaluation on the final zhat.
# ----------------------------------------------------------------------------------------------
@torch.no_grad()
def whole_image_eval(P, x, zhat, dtype=torch.float32):
    """Evaluates the three nets on every block at once, with zhat zero-padded by LRU and valid convs.
    Returns symbols (n,Hb,Wb,M) int32, indexes (n,Hb,Wb,M) int32, xhat (n,3B^2,Hb,Wb) (clamped),
    y (n,M,Hb,Wb), ksi (n,2M,Hb,Wb)."""
    r = P.LRU
    cv = lambda t: t.to(dtype)
    x, zhat = cv(x), cv(zhat)
    zp = F.pad(zhat, (r, r, r, r))
    zp1 = zp if r == 1 else zp[:, :, r - 1:zp.shape[2] - r + 1, r - 1:zp.shape[3] - r + 1]

    def conv(t, p):
        return F.conv2d(t, cv(p[0]), cv(p[1]))

    def gdn(t, p, inverse):
        beta, gamma = cv(p[0]), cv(p[1])
        C = t.shape[1]
        norm = F.conv2d(t ** 2, gamma.reshape(C, C, 1, 1), beta)
        return t * (torch.sqrt(norm) if inverse else torch.rsqrt(norm))

    def chain(t, layers, inverse):
        for i, p in enumerate(layers):
            t = gdn(t, p, inverse) if i % 2 == 0 else conv(t, p)
        return t

    y = chain(conv(x, P.f1) + conv(zp1, P.f2), P.f3, False)
    z = F.leaky_relu(conv(zp, P.e[0]))
    z = F.leaky_relu(conv(z, P.e[1]))
    z = F.leaky_relu(conv(z, P.e[2]))
    ksi = conv(z, P.e[3])
    scales, means = ksi.chunk(2, dim=1)
    s = torch.max(scales.float(), torch.tensor([SCALES_MIN]))
    idx = torch.full(s.shape, SCALES_LEVELS - 1, dtype=torch.int32)
    for t in P.scale_table[:-1]:
        idx -= (s <= t).int()
    sym = torch.round(y - means).int()
    y_qnt = sym.to(dtype) + means
    xhat = chain(conv(y_qnt, P.i1) + conv(zp1, P.i2), P.i3, True).clamp(-0.5, 0.5)
    return (sym.permute(0, 2, 3, 1).contiguous(), idx.permute(0, 2, 3, 1).contiguous(), xhat, y, ksi)


@torch.no_grad()
def forward_open_loop(P, zhat, x, dtype=torch.float32):
    """The reference's model.forward(zhat, x) in eval mode (NET:90-106): ONE whole-image pass of the three nets with
    every masked conv zero-padded by k//2 (NET:266-302), i.e. no closed loop -- zhat is whatever context the caller
    supplies (ACL training-set regeneration feeds the previous iteration's reconstructions, AGENT:643-684).
    Differs from whole_image_eval only for KS[1] = 3: there the second entropy layer sees ZEROS outside the image,
    not the hidden-map ring the windowed codec path produces.
    Returns xhat (n,3B^2,Hb,Wb) NOT clamped, self_informations (n,M,Hb,Wb), symbols (n,Hb,Wb,M) int32."""
    cv = lambda t: t.to(dtype)
    x, zhat = cv(x), cv(zhat)

    def conv(t, p):
        k = p[0].shape[-1]
        return F.conv2d(t, cv(p[0]), cv(p[1]), padding=k // 2)

    def gdn(t, p, inverse):
        beta, gamma = cv(p[0]), cv(p[1])
        C = t.shape[1]
        norm = F.conv2d(t ** 2, gamma.reshape(C, C, 1, 1), beta)
        return t * (torch.sqrt(norm) if inverse else torch.rsqrt(norm))

    def chain(t, layers, inverse):
        for i, p in enumerate(layers):
            t = gdn(t, p, inverse) if i % 2 == 0 else conv(t, p)
        return t

    y = chain(conv(x, P.f1) + conv(zhat, P.f2), P.f3, False)
    z = F.leaky_relu(conv(zhat, P.e[0]))
    z = F.leaky_relu(conv(z, P.e[1]))
    z = F.leaky_relu(conv(z, P.e[2]))
    ksi = conv(z, P.e[3])
    scales, means = ksi.chunk(2, dim=1)
    sym = torch.round(y - means)
    y_qnt = sym + means
    xhat = chain(conv(y_qnt, P.i1) + conv(zhat, P.i2), P.i3, True)
    info = self_information(sym.int(), scales)
    return xhat, info, sym.int().permute(0, 2, 3, 1).contiguous()


def self_information(sym, scales, likelihood_bound: float = 1e-9):
    """-log2 likelihood of quantised latents: ENT:615-647 (eval mode: values = |round(y - mean)|, scales lower-bounded
    at 0.11, likelihood lower-bounded), as summed by validate_recu_reco_fast (AGENT:509-519)."""
    s = torch.max(scales.float(), torch.tensor([SCALES_MIN]))
    v = sym.float().abs()
    c = float(-(2 ** -0.5))
    upper = 0.5 * torch.erfc(c * ((0.5 - v) / s))
    lower = 0.5 * torch.erfc(c * ((-0.5 - v) / s))
    return -torch.log2(torch.max(upper - lower, torch.tensor([likelihood_bound])))


# ----------------------------------------------------------------------------------------------
# layout (AGENT:853-873)
# ----------------------------------------------------------------------------------------------
def arrange_block_pixels_to_channel_dim(x, B):
    """AGENT:853-860: (n,C,H,W) -> (n, C*B*B, H/B, W/B), channel = (v*B+h)*C + c."""
    n, C, H, W = x.shape
    y = torch.empty(n, C * B * B, H // B, W // B, dtype=x.dtype)
    for v in range(B):
        for h in range(B):
            i = (v * B + h) * C
            y[:, i:i + C] = x[:, :, v::B, h::B]
    return y


def arrange_channel_dim_to_block_pixels(y, B):
    """AGENT:863-873."""
    n, CB, Hb, Wb = y.shape
    C = CB // (B * B)
    x = torch.empty(n, C, Hb * B, Wb * B, dtype=y.dtype)
    for v in range(B):
        for h in range(B):
            i = (v * B + h) * C
            x[:, :, v::B, h::B] = y[:, i:i + C]
    return x


# ----------------------------------------------------------------------------------------------
# optional post-processing module (NET:455-476, applied by eval_model at AGENT:604-606)
# ----------------------------------------------------------------------------------------------
@torch.no_grad()
def postprocess(sd, x, clamp=False):
    """BlkBasedPostProcessing.forward: x + F.pad(conv1x1(leaky_relu(conv3x3(x, padding=0))), (1, 1, 1, 1)).
    sd: the module's state_dict (res_net.0.weight/bias, res_net.2.weight/bias); x: (n, 3B^2, Hb, Wb)."""
    res = F.conv2d(F.leaky_relu(F.conv2d(x, sd["res_net.0.weight"], sd["res_net.0.bias"])),
                   sd["res_net.2.weight"], sd["res_net.2.bias"])
    out = x + F.pad(res, (1, 1, 1, 1), "constant", 0) if min(x.shape[2], x.shape[3]) >= 3 else x.clone()
    return out.clamp(-0.5, 0.5) if clamp else out
