"""CPU tests of the single-large-image band pipeline (lbic_b200/band.py): the halo schedule, the lane-container
assembly, and -- over world_size 2 and 3 gloo processes -- the exchange itself, with a toy per-block recurrence that has
the codec's dependency pattern (left, upper-left, up, upper-right) standing in for the GPU step: N ranks in lockstep must
reproduce the single-process result exactly."""
import os
import socket
import struct

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class ToyEngine:
    """z(v,h) = (1 + 3 L + 5 UL + 7 U + 11 UR + x(v,h)) mod 251, channel-wise: exact in fp32, same causal taps as the codec
    (masked_conv2d.py:12-17 mask 'A'), same step structure as lbic_band_step."""
    C = 4

    def begin(self, x, n, Hb, Wb, streams=None, lens=None):
        self.Hb, self.Wb = Hb, Wb
        self.x = x if x is not None else torch.zeros(n, self.C, Hb, Wb)
        self.z = torch.zeros(n, Hb, Wb, self.C)
        return self.z

    def step(self, t, v0, v1):
        z, Hb, Wb = self.z, self.Hb, self.Wb
        g = lambda v, h: z[:, v, h] if (0 <= v < Hb and 0 <= h < Wb) else torch.zeros(z.shape[0], self.C)
        for v in range(max(v0, 0), v1):
            h = t - 2 * v
            if 0 <= h < Wb:
                val = 1 + 3 * g(v, h - 1) + 5 * g(v - 1, h - 1) + 7 * g(v - 1, h) + 11 * g(v - 1, h + 1) + self.x[:, :, v, h]
                z[:, v, h] = torch.remainder(val, 251.0)

    def end(self, v0, v1, want_lanes):
        rows = self.z[:, v0:v1].clone()
        lanes = [bytes(int(q) for q in rows[0, r, :, 0]) for r in range(v1 - v0)] if want_lanes else None
        return rows, lanes


def _single(Hb, Wb):
    x = (torch.arange(ToyEngine.C * Hb * Wb, dtype=torch.float32).reshape(1, ToyEngine.C, Hb, Wb) * 13) % 17
    e = ToyEngine()
    e.begin(x, 1, Hb, Wb)
    for t in range(Wb + 2 * (Hb - 1)):
        e.step(t, 0, Hb)
    return x, e.z


def _worker(rank, world, port, Hb, Wb, q):
    import sys
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import lbic_b200  # noqa: F401
    from lbic_b200 import band
    x, want = _single(Hb, Wb)
    blob, zhat = band.compress_band(None, x, engine=ToyEngine())
    ok = True
    if rank == 0:
        ok = bool(torch.equal(zhat, want.permute(0, 3, 1, 2)))
        lanes = [bytes(int(v) for v in want[0, r, :, 0]) for r in range(Hb)]
        ok = ok and blob == band.pack_lane_container(lanes)
    q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,Hb,Wb", [(2, 6, 9), (3, 7, 5), (2, 2, 1), (3, 9, 20)])
def test_band_exchange_reproduces_single_process(world, Hb, Wb):
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, Hb, Wb, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res), res


def test_halo_schedule_is_consistent():
    """What rank g sends before step t is what rank g+1 receives, and every block a step reads has arrived."""
    import lbic_b200  # noqa: F401
    from lbic_b200 import band
    for Hb, Wb, world in [(6, 9, 2), (16, 16, 4), (5, 3, 5), (64, 96, 8)]:
        bands = [band.band_rows(Hb, world, r) for r in range(world)]
        assert bands[0][0] == 0 and bands[-1][1] == Hb and all(a[1] == b[0] for a, b in zip(bands, bands[1:]))
        have = [set() for _ in range(world)]                       # blocks of foreign rows each rank has received
        for t in range(Wb + 2 * (Hb - 1)):
            for r, (v0, v1) in enumerate(bands):
                recv_h, send_h = band.halo_columns(t, v0, v1, Hb, Wb)
                if r + 1 < world:
                    assert send_h == band.halo_columns(t, bands[r + 1][0], bands[r + 1][1], Hb, Wb)[0]
                if recv_h is not None:
                    have[r].add((v0 - 1, recv_h))
            for r, (v0, v1) in enumerate(bands):
                h = t - 2 * v0                                      # the band's first row reads row v0-1
                if v0 > 0 and 0 <= h < Wb:
                    for hh in (h - 1, h, h + 1):
                        if 0 <= hh < Wb:
                            assert (v0 - 1, hh) in have[r], (Hb, Wb, world, t, r, hh)


def test_lane_container_layout():
    import lbic_b200  # noqa: F401
    from lbic_b200 import band
    blob = band.pack_lane_container([b"abcd", b"", b"12345678"])
    assert blob[:4] == b"LBML" and struct.unpack_from("<IIII", blob, 4) == (3, 4, 0, 8) and blob[20:] == b"abcd12345678"
