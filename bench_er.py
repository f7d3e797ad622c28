#!/usr/bin/env python
"""bench_er.py -- secondary measurement: the expressed-region path (BASELINE.json configs[4], scaled)

    filterData(filter="mean", cutoff=5)  ->  findRegions(meanCoverage)  ->  regionMatrix (coverageMatrix)

through the reference-facing calls with host Rle input (R/regionMatrix.R:170-283).  Not the headline
metric (bench.py is); it puts numbers on SURVEY section 8 rows a1 / a11 (K1 filter kernels, K6
region sums) against the measured HBM peak.  Prints one JSON line.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def full_rle(row, pos_len, pos_first):
    """Rle (values, lengths) over the whole chromosome of one sample: coverage runs inside the kept
    islands, zero runs in the gaps between them."""
    vals, lens = [], []
    o = 0
    v = bool(pos_first)
    for L_ in pos_len:
        L_ = int(L_)
        if L_ == 0:
            v = not v
            continue
        if v:
            x = row[o:o + L_]
            brk = np.flatnonzero(x[1:] != x[:-1]) + 1
            st = np.concatenate(([0], brk))
            vals.append(x[st])
            lens.append(np.diff(np.concatenate((st, [L_]))))
            o += L_
        else:
            vals.append(np.zeros(1, dtype=row.dtype))
            lens.append(np.array([L_]))
        v = not v
    vals, lens = np.concatenate(vals), np.concatenate(lens)
    keep = np.concatenate(([True], vals[1:] != vals[:-1]))  # merge equal neighbours (island edge == 0)
    grp = np.cumsum(keep) - 1
    return vals[keep].astype(np.int32), np.bincount(grp, weights=lens).astype(np.int32)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=100)
    ap.add_argument("--chr-len", type=int, default=24_000_000)
    ap.add_argument("--kept", type=int, default=2_400_000)
    ap.add_argument("--steps", type=int, default=5)
    args = ap.parse_args()
    import torch

    import bench
    import derfinder_b200 as dfb
    from derfinder_b200 import _lib as L

    dev = torch.device("cuda", 0)
    cfg = dict(chr_len=args.chr_len, N=args.kept, n=args.samples, extra=0, B=1, alpha=0.05)
    work = bench.make_workload(cfg, 20140923 + 5, torch=torch, device=dev)
    cov = work["cov"].cpu().numpy()
    cols = [full_rle(cov[s], work["pos_len"], work["pos_first"]) for s in range(args.samples)]
    rle_bytes = sum(v.nbytes + l.nbytes for v, l in cols)
    eng = dfb.Engine(0)
    eng.enable_timing(True)
    times = []
    res = None
    for it in range(args.steps + 1):
        eng.reset_counters()
        t0 = time.perf_counter()
        res = dfb.regionMatrix({"chr21": cols}, cutoff=5, L_=76, returnBP=False, engine=eng)["chr21"]
        dt = time.perf_counter() - t0
        if it:
            times.append(dt)
    k = {name: eng.kernel_time(idx) for name, idx in (("prep", L.K_PREP), ("regions", L.K_REGIONS), ("regmat", L.K_REGMAT))}
    R = len(res["regions"]["start"])
    width = (res["regions"]["indexEnd"] - res["regions"]["indexStart"] + 1).astype(np.int64)
    N = int(width.sum())  # every kept base lies in a region (mean > cutoff on all of them)
    n = args.samples
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = peaks.get("hbm_gbs", 6650.0)
    k1_bytes = rle_bytes + 8.0 * n * N + 8.0 * N + args.chr_len / 8.0  # Rle in, normalised double matrix + mean + mask out
    k6_bytes = 8.0 * n * N + 8.0 * R * n
    line = {"metric": "expressed_region_path_seconds", "value": float(np.median(times)), "unit": "s",
            "higher_is_better": False, "data": "synthetic",
            "config": {"workload": "ER path: filterData(mean>5) + findRegions + regionMatrix", "samples": n,
                       "chr_len": args.chr_len, "kept_bases": N, "regions": R, "rle_input_bytes": rle_bytes},
            "kernels_ms": {kk: vv[0] for kk, vv in k.items()}, "launches": {kk: vv[1] for kk, vv in k.items()},
            "roofline": {"K1_filter": {"bound": "hbm", "achieved": k1_bytes / (k["prep"][0] * 1e-3) / 1e9, "peak": hbm,
                                       "unit": "GB/s", "frac": k1_bytes / (k["prep"][0] * 1e-3) / 1e9 / hbm,
                                       "algorithmic_bytes": k1_bytes},
                         "K6_region_sums": {"bound": "hbm", "achieved": k6_bytes / (k["regmat"][0] * 1e-3) / 1e9,
                                            "peak": hbm, "unit": "GB/s",
                                            "frac": k6_bytes / (k["regmat"][0] * 1e-3) / 1e9 / hbm,
                                            "algorithmic_bytes": k6_bytes}}}
    print(json.dumps(line))
    eng.close()


if __name__ == "__main__":
    main()
