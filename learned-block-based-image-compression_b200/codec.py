"""Image-level codec around the block model: what eval_model does around compress / decompress
(agents/blkbsdimgcomp_agent.py:578-619), plus a self-describing container.

* replicate-pad right / bottom to a multiple of the block size (AGENT:583-586), shift to [-0.5, 0.5], space -> depth;
* compress -> bitstream; the container adds what the reference passes out of band (H, W, B, KS, N, M, lanes) so that
  a stream can be decoded on its own;
* decompress -> depth -> space -> crop to H x W.  The reference crops the block-domain tensor by PIXEL counts
  (AGENT:594, `F.pad(xhat_enc, (0, -padding_right, 0, -padding_bottom))` on the (1, 3B^2, Hb, Wb) tensor), which drops
  whole block columns; here the crop is applied after depth -> space, where it belongs;
* PSNR as AGENT:617 (`-10 log10(mse)` on the [-0.5, 0.5] images), bpp as AGENT:609, MS-SSIM as AGENT:618.

MS-SSIM: the reference imports `pytorch_msssim.ms_ssim`, an un-vendored dependency that is not installable offline, so
`ms_ssim` below restates its published algorithm (11-tap Gaussian, sigma 1.5, five scales, 2x2 average pooling) and is
PARITY UNPINNED against that package.
"""
from __future__ import annotations

import struct
import zlib

import numpy as np
import torch
import torch.nn.functional as F

from .layout import arrange_block_pixels_to_channel_dim, arrange_channel_dim_to_block_pixels

MAGIC = b"LBIC"
VERSION = 1
# magic | version u8 | B u8 | KS u8 x4 | lanes u8 (1 = reference stream, 0 = one lane per block row) | reserved u8
# | N u16 | M u16 | H u32 | W u32 | payload bytes u32 | crc32(payload) u32
_HEADER = struct.Struct("<4sBB4BBBHHIIII")


def pack_container(payload: bytes, *, H: int, W: int, B: int, KS, N: int, M: int, lanes: int) -> bytes:
    ks = [int(k) for k in KS]
    if len(ks) != 4:
        raise ValueError("KS must have four entries")
    head = _HEADER.pack(MAGIC, VERSION, int(B), ks[0], ks[1], ks[2], ks[3], int(lanes), 0, int(N), int(M), int(H), int(W),
                        len(payload), zlib.crc32(payload) & 0xFFFFFFFF)
    return head + payload


def unpack_container(blob: bytes):
    """-> (meta dict, payload bytes); raises ValueError on a foreign, truncated or corrupt container."""
    if len(blob) < _HEADER.size:
        raise ValueError("container shorter than its header")
    magic, ver, B, k0, k1, k2, k3, lanes, _r, N, M, H, W, nbytes, crc = _HEADER.unpack_from(blob, 0)
    if magic != MAGIC:
        raise ValueError("not an LBIC container")
    if ver != VERSION:
        raise ValueError(f"unsupported container version {ver}")
    payload = blob[_HEADER.size:_HEADER.size + nbytes]
    if len(payload) != nbytes:
        raise ValueError("container payload truncated")
    if zlib.crc32(payload) & 0xFFFFFFFF != crc:
        raise ValueError("container payload corrupt (crc mismatch)")
    return dict(H=H, W=W, B=B, KS=[k0, k1, k2, k3], N=N, M=M, lanes=lanes), payload


def pad_to_blocks(x: torch.Tensor, B: int) -> torch.Tensor:
    """(n, 3, H, W) -> replicate-padded on the right / bottom to multiples of B (AGENT:583-586)."""
    h, w = x.shape[2], x.shape[3]
    nh, nw = (h + B - 1) // B * B, (w + B - 1) // B * B
    if nh == h and nw == w:
        return x
    return F.pad(x, (0, nw - w, 0, nh - h), mode="replicate")


def psnr(x: torch.Tensor, y: torch.Tensor) -> float:
    """-10 log10(mse) for images of unit range (AGENT:610-617)."""
    return float(-10.0 * torch.log10(F.mse_loss(x.float(), y.float())))


def _gauss_window(size: int = 11, sigma: float = 1.5, device=None):
    c = torch.arange(size, dtype=torch.float32, device=device) - size // 2
    g = torch.exp(-(c ** 2) / (2 * sigma ** 2))
    return g / g.sum()


def _filter(x, win):
    C = x.shape[1]
    k = win.numel()
    x = F.conv2d(x, win.view(1, 1, k, 1).expand(C, 1, k, 1), groups=C)
    return F.conv2d(x, win.view(1, 1, 1, k).expand(C, 1, 1, k), groups=C)


def _ssim_cs(x, y, win, data_range):
    c1, c2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    mu1, mu2 = _filter(x, win), _filter(y, win)
    s1 = _filter(x * x, win) - mu1 * mu1
    s2 = _filter(y * y, win) - mu2 * mu2
    s12 = _filter(x * y, win) - mu1 * mu2
    cs = (2 * s12 + c2) / (s1 + s2 + c2)
    ssim = ((2 * mu1 * mu2 + c1) / (mu1 * mu1 + mu2 * mu2 + c1)) * cs
    return ssim.flatten(2).mean(-1), cs.flatten(2).mean(-1)


def ms_ssim(x: torch.Tensor, y: torch.Tensor, data_range: float = 1.0) -> torch.Tensor:
    """Multi-scale SSIM of (n, C, H, W) images, mean over channels and batch (the default reduction of
    pytorch_msssim.ms_ssim).  The smaller side must exceed (11 - 1) * 2^4 = 160 pixels."""
    if x.shape != y.shape or x.dim() != 4:
        raise ValueError("ms_ssim expects two (n, C, H, W) tensors of equal shape")
    if min(x.shape[2], x.shape[3]) <= 160:
        raise ValueError("image too small for five-scale MS-SSIM (smaller side must exceed 160)")
    x, y = x.float(), y.float()
    win = _gauss_window(device=x.device)
    weights = torch.tensor([0.0448, 0.2856, 0.3001, 0.2363, 0.1333], device=x.device)
    mcs = []
    for i in range(5):
        ssim, cs = _ssim_cs(x, y, win, data_range)
        if i < 4:
            mcs.append(torch.relu(cs))
            pad = [s % 2 for s in x.shape[2:]]
            x = F.avg_pool2d(x, 2, padding=pad)
            y = F.avg_pool2d(y, 2, padding=pad)
    vals = torch.stack(mcs + [torch.relu(ssim)], dim=0)                 # (5, n, C)
    return torch.prod(vals ** weights.view(-1, 1, 1), dim=0).mean()


class ImageCodec:
    """compress / decompress whole RGB images with a loaded `BlockBasedImgCompLossyNetv9`."""

    def __init__(self, model, lanes: int = 1):
        self.model = model
        self.lanes = int(lanes)
        cfg = model.config
        self.B, self.KS, self.N, self.M = int(cfg.block_size), [int(k) for k in cfg.KS], int(cfg.N), int(cfg.M)
        self.LRU = [sum(k // 2 for k in self.KS)] * 3

    # ---- tensors ---------------------------------------------------------------------------------
    def _to_blocks(self, img: torch.Tensor):
        """img (n, 3, H, W) in [0, 1] on the model's device -> padded block tensor in [-0.5, 0.5]."""
        x = pad_to_blocks(img.float() - 0.5, self.B)
        return arrange_block_pixels_to_channel_dim(x, self.B).contiguous()

    def _from_blocks(self, z: torch.Tensor, H: int, W: int):
        return (arrange_channel_dim_to_block_pixels(z, self.B)[:, :, :H, :W] + 0.5).clamp_(0.0, 1.0)

    def encode(self, img) -> bytes:
        """img: (3, H, W) / (1, 3, H, W) float tensor in [0, 1], or an (H, W, 3) uint8 array -> container bytes."""
        t = self._as_tensor(img)
        H, W = int(t.shape[2]), int(t.shape[3])
        x = self._to_blocks(t)
        if self.lanes == 1:
            payload, _ = self.model.compress(x, self.LRU, self.M)
        else:
            payload = self.model.compress_batch(x, lanes=0)[0][0]
        return pack_container(payload, H=H, W=W, B=self.B, KS=self.KS, N=self.N, M=self.M, lanes=self.lanes)

    def decode(self, blob: bytes) -> torch.Tensor:
        """container bytes -> (1, 3, H, W) float tensor in [0, 1] on the model's device."""
        meta, payload = unpack_container(blob)
        if (meta["B"], meta["KS"], meta["N"], meta["M"]) != (self.B, self.KS, self.N, self.M):
            raise ValueError(f"container was written by a different model configuration: {meta}")
        Hb, Wb = (meta["H"] + self.B - 1) // self.B, (meta["W"] + self.B - 1) // self.B
        shape = (1, 3 * self.B * self.B, Hb, Wb)
        if meta["lanes"] == 1:
            z = self.model.decompress(payload, self.LRU, shape, self.M, self.model._device)
        else:
            z = self.model.decompress_batch([payload], shape, lanes=0)
        return self._from_blocks(z, meta["H"], meta["W"])

    def evaluate(self, img) -> dict:
        """One image through encode + decode with the figures eval_model logs (AGENT:609-619)."""
        t = self._as_tensor(img)
        blob = self.encode(t)
        rec = self.decode(blob)
        meta, payload = unpack_container(blob)
        # the figures of AGENT:611-619 on the GPU (lbic_image_metrics), on the [-0.5, 0.5] images the agent compares
        big = min(t.shape[2], t.shape[3]) > 160
        q = self.model.image_metrics((t - 0.5).contiguous(), (rec - 0.5).contiguous(), msssim=big)
        out = dict(bytes=len(payload), container_bytes=len(blob), bpp=8.0 * len(payload) / (t.shape[2] * t.shape[3]),
                   mse=float(q["mse"][0]), psnr=float(q["psnr"][0]))
        if big:
            m = float(q["msssim"][0])
            out.update(msssim=m, msssim_db=float(-10.0 * np.log10(max(1.0 - m, 1e-12))))
        return out

    # ---- files -----------------------------------------------------------------------------------
    def encode_file(self, path: str) -> bytes:
        from PIL import Image
        with Image.open(path) as im:
            return self.encode(np.asarray(im.convert("RGB")))

    def decode_to_file(self, blob: bytes, path: str):
        from PIL import Image
        rec = self.decode(blob)[0]
        arr = (rec * 255.0).round().clamp_(0, 255).to(torch.uint8).permute(1, 2, 0).cpu().numpy()
        Image.fromarray(arr, "RGB").save(path)

    def _as_tensor(self, img) -> torch.Tensor:
        if isinstance(img, np.ndarray):
            if img.ndim != 3 or img.shape[2] != 3 or img.dtype != np.uint8:
                raise ValueError("arrays must be (H, W, 3) uint8")
            t = torch.from_numpy(np.array(img, copy=True)).permute(2, 0, 1).float().div_(255.0)
        else:
            t = img
        if t.dim() == 3:
            t = t[None]
        if t.dim() != 4 or t.shape[0] != 1 or t.shape[1] != 3:
            raise ValueError("expected one RGB image: (3, H, W) or (1, 3, H, W)")
        return t.to(self.model._device)
