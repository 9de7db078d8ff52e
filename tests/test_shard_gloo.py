"""CPU test of the N>1 host logic: world_size-2 gloo processes shard images and gather variable-length
bitstreams in global image order (the data path itself has no collective)."""
import os
import socket

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


def _fake_stream(i):
    return bytes((i * 7 + k) % 251 for k in range(8 + 4 * (i % 5)))


def _worker(rank, world, port, n_images, q):
    import sys
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lbic_b200 import shard
    start, count = shard.image_shard(n_images, world, rank)
    local = [_fake_stream(i) for i in range(start, start + count)]
    everything = shard.gather_bitstreams(local)
    rows = shard.gather_rows(torch.arange(start, start + count, dtype=torch.float32).reshape(-1, 1).repeat(1, 3))
    q.put((rank, start, count, everything == [_fake_stream(i) for i in range(n_images)],
           rows[:, 0].tolist() == [float(i) for i in range(n_images)]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_images", [5, 8, 1])
def test_two_rank_shard_and_gather(n_images):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_images, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sum(r[2] for r in res) == n_images and res[0][1] == 0 and res[1][1] == res[0][2]
    assert all(r[3] and r[4] for r in res)


def test_image_shard_covers_everything():
    import lbic_b200  # noqa: F401
    from lbic_b200 import shard
    for n in (0, 1, 7, 24, 1024):
        for w in (1, 2, 4, 8):
            seen = []
            for r in range(w):
                s, c = shard.image_shard(n, w, r)
                seen += list(range(s, s + c))
            assert seen == list(range(n))
