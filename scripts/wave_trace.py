"""Reads a wave-kernel trace (LBIC_WAVE_TRACE, gemm_wave.cu) and prints, per step and list entry, where the time goes:
    wait   first tile's inputs ready, relative to the step's first event
    tma    inputs ready -> first pipeline stage full          (median over the entry's tiles)
    main   first stage full -> accumulator complete            (operand streaming + MMAs)
    epi    accumulator complete -> stores issued
    pub    stores issued -> counter bumped (store drain + gpu-scope release)
    span   first input ready -> last tile published
    python scripts/wave_trace.py gpurun_out/wave_trace.txt
"""
import statistics
import sys
from collections import defaultdict

NAMES = {0: "GEMM", 1: "GATHER", 2: "RANS", 3: "GATHER_EXT", 4: "GATHER5"}
LAYERS = ("E0", "E1", "E2", "E3", "F0", "G0", "F1", "G1", "F2", "G2", "F3", "D0", "IG0", "D1", "IG1", "D2", "IG2", "D3")


def main(path):
    steps = defaultdict(lambda: defaultdict(list))
    for line in open(path):
        if line.startswith("#"):
            print(line.strip())
            continue
        a, b = line.split("|")
        s, j, oi, kind, layer, rb, nt, cta = (int(v) for v in a.split())
        t = [int(v) for v in b.split()]
        steps[s][oi].append(dict(kind=kind, layer=layer, rb=rb, nt=nt, cta=cta, t=t))
    med = lambda v: statistics.median(v) if v else 0
    for s in sorted(steps):
        ents = steps[s]
        t0 = min(min(x["t"][0] for x in tiles if x["t"][0]) for tiles in ents.values())
        t_end = max(max(x["t"][6] for x in tiles) for tiles in ents.values())
        print(f"step {s}: {(t_end - t0) / 1e3:.1f} us, {sum(len(v) for v in ents.values())} tiles")
        print(f"  {'entry':10s} {'tiles':>5s} {'wait':>7s} {'tma':>6s} {'main':>6s} {'epi':>6s} {'pub':>6s} {'span':>7s} {'done@':>7s}")
        for oi in sorted(ents):
            tiles = ents[oi]
            k = tiles[0]["kind"]
            name = LAYERS[tiles[0]["layer"]] if k == 0 else NAMES[k]
            ready = [x["t"][0] for x in tiles if x["t"][0]]
            if k == 0:
                tma = med([x["t"][2] - x["t"][0] for x in tiles])
                main_ = med([x["t"][4] - x["t"][2] for x in tiles])
                epi = med([x["t"][5] - x["t"][4] for x in tiles])
            else:
                tma, main_ = 0, 0
                epi = med([x["t"][5] - x["t"][0] for x in tiles])
            pub = med([x["t"][6] - x["t"][5] for x in tiles])
            span = max(x["t"][6] for x in tiles) - min(ready)
            print(f"  {name:10s} {len(tiles):5d} {(min(ready) - t0) / 1e3:7.2f} {tma / 1e3:6.2f} {main_ / 1e3:6.2f} {epi / 1e3:6.2f} "
                  f"{pub / 1e3:6.2f} {span / 1e3:7.2f} {(max(x['t'][6] for x in tiles) - t0) / 1e3:7.2f}")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/wave_trace.txt")
