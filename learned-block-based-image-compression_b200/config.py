"""Reads the reference's own config JSONs (configs/*.json; utils/config.py:50-67 turns them into an
EasyDict).  Only the keys the codec path consumes are required: block_size, KS, N, M
(graphs/models/BlockBasedImgCompLossy_net.py:262-302); the agent-level keys (net_version, mode,
cuda, gpu_device, seed, val_batch_size, lambda_ ...) are carried through untouched."""
from __future__ import annotations

import json
from types import SimpleNamespace

BUILTIN = {
    # the four shipped reference configs (configs/blkbsdimgcomp_*.json), codec keys only
    "B8_lowrate": dict(block_size=8, KS=[3, 1, 1, 1], N=768, M=96),
    "B4_highrate": dict(block_size=4, KS=[3, 3, 1, 1], N=512, M=96),
    "B8_highrate": dict(block_size=8, KS=[3, 3, 1, 1], N=1152, M=128),
    "B16_lowrate": dict(block_size=16, KS=[3, 1, 1, 1], N=1280, M=192),
}


def load_config(src) -> SimpleNamespace:
    """src: path to a reference JSON, a dict, a namespace, or one of the BUILTIN names."""
    if isinstance(src, str) and src in BUILTIN:
        d = dict(BUILTIN[src], net_version="v9", seed=1337)
    elif isinstance(src, str):
        with open(src) as f:
            d = json.load(f)
    elif isinstance(src, dict):
        d = dict(src)
    else:
        d = dict(vars(src))
    for k in ("block_size", "KS", "N", "M"):
        if k not in d:
            raise KeyError(f"config is missing required key '{k}'")
    if d.get("net_version", "v9") != "v9":
        raise ValueError("only net_version 'v9' is supported (all shipped reference configs use it)")
    if int(d["KS"][0]) != 3 or int(d["KS"][1]) not in (1, 3) or list(map(int, d["KS"][2:])) != [1, 1]:
        raise ValueError(f"unsupported KS {d['KS']}: the reference configs use [3,1,1,1] or [3,3,1,1]")
    return SimpleNamespace(**d)
