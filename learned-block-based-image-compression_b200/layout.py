"""arrange_block_pixels_to_channel_dim / arrange_channel_dim_to_block_pixels
(agents/blkbsdimgcomp_agent.py:853-873) as CUDA kernels behind the C ABI.  Same names, same
argument order (x, B, dev) as the reference helpers."""
from __future__ import annotations

import torch

from . import _lib


def _stream_ptr(t: torch.Tensor):
    return torch.cuda.current_stream(t.device).cuda_stream


def arrange_block_pixels_to_channel_dim(x: torch.Tensor, B: int, dev=None) -> torch.Tensor:
    """(n, C, H, W) -> (n, C*B*B, H/B, W/B) with channel index (v*B+h)*C + c."""
    if not x.is_cuda:
        raise RuntimeError("lbic_b200 layout kernels need a CUDA tensor (no CPU fallback)")
    x = x.contiguous().float()
    n, C, H, W = x.shape
    assert H % B == 0 and W % B == 0, "H and W must be multiples of B (pad first, AGENT:583-586)"
    y = torch.empty(n, C * B * B, H // B, W // B, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().lbic_space_to_depth(x.data_ptr(), y.data_ptr(), n, C, H // B, W // B, B, _stream_ptr(x)))
    return y


def arrange_channel_dim_to_block_pixels(y: torch.Tensor, B: int, dev=None) -> torch.Tensor:
    """(n, C*B*B, Hb, Wb) -> (n, C, Hb*B, Wb*B)."""
    if not y.is_cuda:
        raise RuntimeError("lbic_b200 layout kernels need a CUDA tensor (no CPU fallback)")
    y = y.contiguous().float()
    n, CB, Hb, Wb = y.shape
    C = CB // (B * B)
    x = torch.empty(n, C, Hb * B, Wb * B, device=y.device, dtype=torch.float32)
    with torch.cuda.device(y.device):
        _lib.check(_lib.lib().lbic_depth_to_space(y.data_ptr(), x.data_ptr(), n, C, Hb, Wb, B, _stream_ptr(y)))
    return x
