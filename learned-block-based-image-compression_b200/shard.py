"""Multi-GPU plumbing for the codec path: images are independent units (eval_model loops them one by one,
agents/blkbsdimgcomp_agent.py:578; no cross-image state), so N ranks take contiguous image ranges and the
data path needs NO collective.  Collectives are used only to gather results: per-image stream lengths, then
the variable-length bitstreams (and optionally reconstructions).  Works with NCCL (CUDA tensors) and gloo."""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def image_shard(n_images: int, world_size: int, rank: int):
    """Contiguous range [start, start+count) of images owned by `rank`; the remainder goes to the first ranks."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, rem = divmod(int(n_images), int(world_size))
    count = base + (1 if rank < rem else 0)
    start = rank * base + min(rank, rem)
    return start, count


def gather_bitstreams(local_strings, device="cpu", group=None):
    """All ranks contribute their images' bitstreams (list[bytes], in local image order); every rank returns
    the full list in global image order.  Two collectives: all_gather of counts+lengths, all_gather of a
    padded uint8 buffer."""
    if not dist.is_initialized():
        return list(local_strings)
    world = dist.get_world_size(group)
    n_local = torch.tensor([len(local_strings)], dtype=torch.int64, device=device)
    counts = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(counts, n_local, group=group)
    counts = [int(c.item()) for c in counts]
    max_n = max(counts) if counts else 0
    lens = torch.zeros(max(max_n, 1), dtype=torch.int64, device=device)
    for i, s in enumerate(local_strings):
        lens[i] = len(s)
    all_lens = [torch.zeros_like(lens) for _ in range(world)]
    dist.all_gather(all_lens, lens, group=group)
    max_len = max(int(l.max().item()) for l in all_lens)
    buf = np.zeros((max(max_n, 1), max(max_len, 1)), dtype=np.uint8)
    for i, s in enumerate(local_strings):
        buf[i, : len(s)] = np.frombuffer(s, dtype=np.uint8)
    t = torch.from_numpy(buf).to(device)
    all_buf = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(all_buf, t, group=group)
    out = []
    for r in range(world):
        b, l = all_buf[r].cpu().numpy(), all_lens[r].cpu().numpy()
        out += [b[i, : int(l[i])].tobytes() for i in range(counts[r])]
    return out


def gather_rows(local: torch.Tensor, group=None):
    """Concatenates per-rank tensors with possibly different leading sizes (e.g. zhat of each rank's images)."""
    if not dist.is_initialized():
        return local
    world = dist.get_world_size(group)
    n = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    ns = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(ns, n, group=group)
    ns = [int(v.item()) for v in ns]
    pad = torch.zeros((max(ns),) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[:k] for p, k in zip(parts, ns)], dim=0)
