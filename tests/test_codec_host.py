"""CPU tests of the image-level host logic: container, padding, metrics."""
import numpy as np
import pytest
import torch

from lbic_b200 import codec


def test_container_roundtrip_and_rejects_damage():
    payload = bytes(range(256)) * 3
    blob = codec.pack_container(payload, H=511, W=770, B=8, KS=[3, 1, 1, 1], N=768, M=96, lanes=1)
    meta, out = codec.unpack_container(blob)
    assert out == payload
    assert meta == dict(H=511, W=770, B=8, KS=[3, 1, 1, 1], N=768, M=96, lanes=1)
    with pytest.raises(ValueError):
        codec.unpack_container(blob[:10])
    with pytest.raises(ValueError):
        codec.unpack_container(b"XXXX" + blob[4:])
    with pytest.raises(ValueError):
        codec.unpack_container(blob[:-1])                       # truncated payload
    damaged = bytearray(blob)
    damaged[-5] ^= 0x40
    with pytest.raises(ValueError):
        codec.unpack_container(bytes(damaged))                  # crc
    assert codec.unpack_container(codec.pack_container(b"", H=8, W=8, B=8, KS=[3, 3, 1, 1], N=512, M=192, lanes=0))[1] == b""


def test_replicate_padding_matches_reference_rule():
    # AGENT:583-586: pad right / bottom with mode='replicate' up to the next multiple of B
    x = torch.arange(2 * 3 * 5 * 7, dtype=torch.float32).reshape(2, 3, 5, 7)
    p = codec.pad_to_blocks(x, 4)
    assert p.shape == (2, 3, 8, 8)
    assert torch.equal(p[:, :, :5, :7], x)
    assert torch.equal(p[:, :, 5:, :7], x[:, :, 4:5, :].expand(-1, -1, 3, -1))
    assert torch.equal(p[:, :, :, 7], p[:, :, :, 6])
    assert codec.pad_to_blocks(p, 4) is p


def test_psnr_and_ms_ssim_properties():
    g = torch.Generator().manual_seed(3)
    low = torch.rand(1, 3, 12, 16, generator=g)
    x = torch.nn.functional.interpolate(low, size=(192, 256), mode="bicubic", align_corners=False).clamp(0, 1)
    assert float(codec.ms_ssim(x, x)) == pytest.approx(1.0, abs=1e-6)
    noisy = (x + 0.05 * torch.randn(x.shape, generator=g)).clamp(0, 1)
    noisier = (x + 0.15 * torch.randn(x.shape, generator=g)).clamp(0, 1)
    a, b = float(codec.ms_ssim(x, noisy)), float(codec.ms_ssim(x, noisier))
    assert 0.0 < b < a < 1.0
    assert float(codec.ms_ssim(noisy, x)) == pytest.approx(a, rel=1e-5)            # symmetric
    assert codec.psnr(x, noisy) > codec.psnr(x, noisier)
    assert codec.psnr(x, x + 0.1) == pytest.approx(20.0, abs=1e-4)                   # mse = 0.01
    with pytest.raises(ValueError):
        codec.ms_ssim(x[:, :, :100], x[:, :, :100])
    # one scale of the same SSIM agrees with a direct evaluation of the formula on a flat patch pair
    f1, f2 = torch.full((1, 1, 176, 176), 0.25), torch.full((1, 1, 176, 176), 0.75)
    c1 = 0.01 ** 2
    expect = (2 * 0.25 * 0.75 + c1) / (0.25 ** 2 + 0.75 ** 2 + c1)                   # zero variance: cs = 1 at every scale
    assert float(codec.ms_ssim(f1, f2)) == pytest.approx(expect ** 0.1333, rel=1e-3)   # fp32 cancellation in E[x^2] - mu^2
