"""Reads an `ncu --set full` report (ncu -i <rep> --page raw --csv) and summarises the launches of one kernel:
duration, DRAM bytes read + written per launch, tensor-pipe activity, L2 hit rate, achieved occupancy, registers.
With --key it also records the DRAM traffic per launch in profiles/ncu_traffic.json, which bench.py reports as
roofline.traffic for that workload.

    ncu -i gpurun_out/r2_flow.ncu-rep --page raw --csv > /tmp/raw.csv
    python scripts/ncu_traffic.py /tmp/raw.csv --kernel gemm_flow --key B8_lowrate:1024:768x512:lanes0 --note "..."
"""
import argparse
import csv
import json
import os
import statistics

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = {"gpu__time_duration.sum": "duration", "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active": "tensor_hmma_pct",
        "lts__t_sector_hit_rate.pct": "l2_hit_pct", "lts__t_bytes.sum": "l2_bytes",
        "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct", "launch__registers_per_thread": "registers",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
        "sm__cycles_active.avg": "sm_cycles_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct"}
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3,
         "second": 1.0, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--kernel", default="")
    ap.add_argument("--key", default="")
    ap.add_argument("--note", default="")
    args = ap.parse_args()
    rows = list(csv.reader(open(args.csv)))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units = rows[hdr], rows[hdr + 1]
    col = {n: i for i, n in enumerate(names)}
    out = []
    for r in rows[hdr + 2:]:
        if len(r) < len(names) or args.kernel not in r[col["Kernel Name"]]:
            continue
        rec = dict(kernel=r[col["Kernel Name"]][:60])
        for metric, key in WANT.items():
            if metric in col:
                try:
                    v = float(r[col[metric]].replace(",", ""))
                except ValueError:
                    continue
                rec[key] = v * SCALE.get(units[col[metric]], 1)
        out.append(rec)
    if not out:
        raise SystemExit("no matching launches")
    med = lambda k: statistics.median([r[k] for r in out if k in r]) if any(k in r for r in out) else None
    summary = dict(launches=len(out), kernel=out[0]["kernel"])
    for k in list(WANT.values()):
        summary[k] = med(k)
    if summary.get("dram_read") is not None:
        summary["dram_bytes_per_launch"] = summary["dram_read"] + summary["dram_write"]
        if summary.get("duration"):
            summary["dram_gb_s"] = summary["dram_bytes_per_launch"] / summary["duration"] / 1e9
    print(json.dumps(summary, indent=1))
    if args.key:
        p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        db = json.load(open(p)) if os.path.exists(p) else {}
        db[args.key] = dict(dram_bytes_per_launch=summary["dram_bytes_per_launch"], kernel=summary["kernel"],
                            duration_s_under_ncu=summary.get("duration"), tensor_pipe_pct=summary.get("tensor_pipe_pct"),
                            l2_hit_pct=summary.get("l2_hit_pct"), note=args.note)
        json.dump(db, open(p, "w"), indent=1)


if __name__ == "__main__":
    main()
