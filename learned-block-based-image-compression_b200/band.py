"""One LARGE image over several GPUs: block-row bands with a per-step halo exchange (BASELINE.json config 5).

The closed loop (graphs/models/BlockBasedImgCompLossy_net.py:339-357 / 420-450) lets block (v, h) read the reconstructed
blocks (v, h-1), (v-1, h-1), (v-1, h), (v-1, h+1) (KS[1] = 1), so a horizontal cut between block rows v1-1 and v1 is
crossed by exactly one new block per wavefront step t = h + 2 v: zhat(v1-1, t - 2 (v1-1)) is produced by the upper rank
in step t and first read by the lower rank in step t+1.  Rank g owns the contiguous rows [v0, v1) = band_rows(...);
every rank walks all steps in lockstep, runs `lbic_band_step` on its own rows and passes that one block down with
torch.distributed send / recv (NCCL over NVLink on a B200 box; gloo in the CPU test).  Nothing else is exchanged on
the data path: each band is entropy-coded as the lanes of its own rows, and the bands' lanes are gathered into the
same 'LBML' container a single GPU writes (so `decompress_batch(lanes=0)` on one GPU decodes what N GPUs encoded and
vice versa).

What this buys: nothing below ~2000 block rows per step.  A step's time is set by the dependent chain of 14 layers
(~150-210 us), not by how many rows it has (a single 8192x8192 image has at most 512), so the ranks spend the same
number of steps at the same per-step latency plus the exchange (DESIGN.md section 7 has the measurements).
"""
from __future__ import annotations

import ctypes
import struct

import numpy as np
import torch
import torch.distributed as dist

from . import _lib, shard

LANE_MAGIC = 0x4C4D424C   # 'LBML'


def band_rows(Hb: int, world: int, rank: int):
    """Contiguous block rows [v0, v1) of `rank` (bands differ by at most one row)."""
    return Hb * rank // world, Hb * (rank + 1) // world


def halo_columns(t: int, v0: int, v1: int, Hb: int, Wb: int):
    """Before step t: (column of the block of row v0-1 to RECEIVE from the rank above or None,
                      column of the block of row v1-1 to SEND to the rank below or None).
    Both are the blocks computed in step t-1, the last ones block row v0 / v1 has not seen yet."""
    def col(v):
        h = (t - 1) - 2 * v
        return h if 0 <= h < Wb else None
    recv = col(v0 - 1) if v0 > 0 else None
    send = col(v1 - 1) if v1 < Hb else None
    return recv, send


def pack_lane_container(lanes):
    """list of per-block-row rANS streams -> the lane container of rans.cu: 'LBML' | n | len[n] | payloads."""
    head = struct.pack("<II", LANE_MAGIC, len(lanes)) + b"".join(struct.pack("<I", len(s)) for s in lanes)
    return head + b"".join(lanes)


class _DevMem:
    """A raw device pointer as a CUDA array (so that torch can wrap library-owned memory without a copy)."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = dict(shape=tuple(shape), typestr="<f4", data=(int(ptr), False), version=3, strides=None)


class LibEngine:
    """The band steps of liblbic_b200 on this rank's GPU."""

    def __init__(self, model):
        self.m = model
        self.L = _lib.lib()

    def begin(self, x, n, Hb, Wb, streams=None, lens=None):
        h = self.m._need()
        self.n, self.Hb, self.Wb = n, Hb, Wb
        self._keep = (x, streams, lens)
        with torch.cuda.device(self.m._device):
            _lib.check(self.L.lbic_band_begin(h, x.data_ptr() if x is not None else None, n, Hb, Wb,
                                              streams.data_ptr() if streams is not None else None,
                                              lens.data_ptr() if lens is not None else None,
                                              streams.shape[1] if streams is not None else 0, self.m._stream()))
        ptr = self.L.lbic_band_zhat(h)
        return torch.as_tensor(_DevMem(ptr, (n, Hb, Wb, self.m.Cin)), device=self.m._device)

    def step(self, t, v0, v1):
        _lib.check(self.L.lbic_band_step(self.m._need(), t, v0, v1, self.m._stream()))

    def end(self, v0, v1, want_lanes):
        m, n, nr = self.m, self.n, v1 - v0
        rows = torch.empty(n, nr, self.Wb, m.Cin, dtype=torch.float32, device=m._device)
        lanes, cap = None, 0
        if want_lanes:
            cap = (8 * self.Wb * m.M + 64 + 3) // 4 * 4
            buf = torch.empty(n * nr, cap, dtype=torch.uint8, device=m._device)
            ln = torch.zeros(n * nr, dtype=torch.int32, device=m._device)
        with torch.cuda.device(m._device):
            _lib.check(self.L.lbic_band_end(m._need(), v0, v1, rows.data_ptr(), buf.data_ptr() if want_lanes else None, cap,
                                            ln.data_ptr() if want_lanes else None, m._stream()))
        m.check_errors()          # synchronises: an overflowing lane buffer or a malformed container raises here
        if want_lanes:
            lh = ln.cpu().numpy().astype(np.int64)
            bh = buf[:, : int(lh.max())].cpu().numpy()
            lanes = [bh[i, : lh[i]].tobytes() for i in range(n * nr)]
        return rows, lanes


def _run_steps(engine, zt, Hb, Wb, v0, v1, rank, world, group):
    """All wavefront steps in lockstep; before each, the halo block of the previous step moves one rank down."""
    for t in range(Wb + 2 * (Hb - 1)):
        recv_h, send_h = halo_columns(t, v0, v1, Hb, Wb)
        ops = []
        if recv_h is not None and rank > 0:
            ops.append(dist.P2POp(dist.irecv, zt[:, v0 - 1, recv_h], rank - 1, group))
        if send_h is not None and rank < world - 1:
            ops.append(dist.P2POp(dist.isend, zt[:, v1 - 1, send_h], rank + 1, group))
        if ops:
            for r in dist.batch_isend_irecv(ops):
                r.wait()
        engine.step(t, v0, v1)


def _world(group):
    if not dist.is_initialized():
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def compress_band(model, x, group=None, engine=None):
    """x: (1, 3B^2, Hb, Wb) on this rank's device, the SAME full image on every rank (each rank reads its own rows).
    Returns (lane-container bytes, zhat (1, 3B^2, Hb, Wb)) on rank 0 and (None, None) elsewhere."""
    rank, world = _world(group)
    n, _, Hb, Wb = x.shape
    if n != 1:
        raise ValueError("band mode splits ONE image; batches are sharded by image (shard.py)")
    if world > Hb:
        raise ValueError("more ranks than block rows")
    eng = engine or LibEngine(model)
    v0, v1 = band_rows(Hb, world, rank)
    zt = eng.begin(x.contiguous(), n, Hb, Wb)
    _run_steps(eng, zt, Hb, Wb, v0, v1, rank, world, group)
    rows, lanes = eng.end(v0, v1, True)
    all_lanes = shard.gather_bitstreams(lanes, device=rows.device if rows.is_cuda else "cpu", group=group)
    full = shard.gather_rows(rows[0], group=group)                         # (Hb, Wb, Cin)
    if rank != 0:
        return None, None
    return pack_lane_container(all_lanes), full.permute(2, 0, 1).unsqueeze(0).contiguous()


def decompress_band(model, blob, xshape, group=None, engine=None):
    """blob: the lane container (every rank passes the same bytes).  Returns zhat on rank 0, None elsewhere."""
    rank, world = _world(group)
    n, _, Hb, Wb = [int(v) for v in xshape]
    if n != 1:
        raise ValueError("band mode splits ONE image")
    eng = engine or LibEngine(model)
    v0, v1 = band_rows(Hb, world, rank)
    dev = model._device if engine is None else "cpu"
    cap = (len(blob) + 3) // 4 * 4
    host = np.zeros((1, cap), dtype=np.uint8)
    host[0, : len(blob)] = np.frombuffer(blob, dtype=np.uint8)
    streams = torch.from_numpy(host).to(dev)
    lens = torch.tensor([len(blob)], dtype=torch.int32, device=dev)
    zt = eng.begin(None, n, Hb, Wb, streams, lens)
    _run_steps(eng, zt, Hb, Wb, v0, v1, rank, world, group)
    rows, _ = eng.end(v0, v1, False)
    full = shard.gather_rows(rows[0], group=group)
    if rank != 0:
        return None
    return full.permute(2, 0, 1).unsqueeze(0).contiguous()
