"""BASELINE.json config 5: ONE large synthetic image (default 8192x8192, B8 N768 M96) encoded + decoded in block-row
bands over N GPUs with a per-step halo exchange (lbic_b200/band.py), against the same image on one GPU.

    python scripts/band_bench.py [--size 8192]                                           (1 GPU: single-GPU numbers only)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           scripts/band_bench.py [--size 8192]
Prints one JSON line on rank 0: band encode / decode ms (CUDA events, max over ranks), the single-GPU per-layer and
wave-kernel times on rank 0, and whether the N-GPU stream and reconstruction are bit-identical to the single-GPU ones."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lbic_b200  # noqa: E402
from lbic_b200 import band, weights  # noqa: E402
from lbic_b200.layout import arrange_block_pixels_to_channel_dim  # noqa: E402
from lbic_b200.net import BlockBasedImgCompLossyNetv9  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=8192)
    ap.add_argument("--width", type=int, default=0)
    ap.add_argument("--config", default="B8_lowrate")
    ap.add_argument("--reps", type=int, default=2)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = lbic_b200.load_config(args.config)
    m = BlockBasedImgCompLossyNetv9(cfg, device=dev)
    m.load_state_dict(weights.synth_state_dict(cfg, 1337))
    m.update(force=True)
    B, H, W = int(cfg.block_size), args.size, args.width or args.size
    g = torch.Generator(device=dev)
    g.manual_seed(4242)                                     # the same image on every rank
    low = torch.rand(1, 3, H // 16, W // 16, generator=g, device=dev)
    img = (torch.nn.functional.interpolate(low, size=(H, W), mode="bicubic", align_corners=False)
           + 0.05 * torch.randn(1, 3, H, W, generator=g, device=dev)).clamp_(0, 1)
    x = arrange_block_pixels_to_channel_dim(img - 0.5, B)
    del img, low
    Hb, Wb = H // B, W // B

    def timed(fn):
        best, out = 1e30, None
        for _ in range(args.reps):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            out = fn()
            b.record()
            torch.cuda.synchronize()
            ms = torch.tensor([a.elapsed_time(b)], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            best = min(best, float(ms.item()))
        return best, out

    band.compress_band(m, x)                                 # warm-up (workspace, NCCL channels)
    ms_enc, (blob, zhat) = timed(lambda: band.compress_band(m, x))
    blob_all = [blob]
    if world > 1:
        dist.broadcast_object_list(blob_all, src=0)
    band.decompress_band(m, blob_all[0], x.shape)
    ms_dec, zdec = timed(lambda: band.decompress_band(m, blob_all[0], x.shape))
    if rank == 0:
        rec = dict(config=args.config, image=f"{W}x{H}", n_gpus=world, block_rows=Hb, steps=Wb + 2 * (Hb - 1),
                   band_encode_ms=round(ms_enc, 2), band_decode_ms=round(ms_dec, 2),
                   band_encode_mpix_s=round(H * W / ms_enc / 1e3, 2), band_decode_mpix_s=round(H * W / ms_dec / 1e3, 2),
                   enc_dec_identical=bool(torch.equal(zdec, zhat)), stream_bytes=len(blob))
        # the same image on ONE GPU through the regular calls
        for wave in (0, 1):
            m.set_option("wave", wave)
            o = m.encode_device(x, lanes=0)
            torch.cuda.synchronize()
            a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            a.record()
            o = m.encode_device(x, lanes=0, out=o)
            b.record()
            z1 = m.decode_device(o.streams, o.lens, 1, Hb, Wb, lanes=0)
            c.record()
            torch.cuda.synchronize()
            rec[f"single_gpu_{'wave' if wave else 'layer'}_encode_ms"] = round(a.elapsed_time(b), 2)
            rec[f"single_gpu_{'wave' if wave else 'layer'}_decode_ms"] = round(b.elapsed_time(c), 2)
        one = m._gather_streams(o)[0]
        rec["stream_identical_to_single_gpu"] = bool(one == blob)
        rec["zhat_identical_to_single_gpu"] = bool(torch.equal(o.zhat, zhat))
        print(json.dumps(rec), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
