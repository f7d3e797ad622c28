#!/usr/bin/env python
"""bench.py -- the derfinder F-stat / region p-value hot path on N B200s (one process per GPU).

    python bench.py --gpus 1 --steps K --warmup W            # our CUDA engine
    python bench.py --impl reference --steps K --warmup W    # the reference's algorithm on host cores

Metric (BASELINE.json): base x permutation F-stats per second = N * (B + 1) / t, where one "step"
is one full pass of the hot path over one synthetic chromosome:
    preprocessCoverage (means + log2(x+32))  ->  calculateStats (observed F)
    ->  calculatePvalues (B permuted-model F scans, thresholding, null regions/areas,
        observed regions + clusters, null-area sort, exact p-values)
`value` times it with the filtered coverage already resident in HBM; `e2e` times the same pass
through the reference-facing calls with HOST buffers (H2D of the coverage -- as integer Rle columns,
the type the reference holds it in -- and D2H of F-stats, regions, nulls and p-values inside the
timed region).

Workload (default `c3`): BASELINE.json configs[2], the largest single-GPU configuration -- synthetic
chr1 (Lc = 249,250,621; 25 M bases pass the filter), 60 samples, 2-group model + sampleDepth + 3
adjustment covariates (p = 6, p0 = 5), 1000 permutations, F cutoff qf(1e-6, 1, 54).  `--workload c2`
is configs[1] (chr21, 12 samples, 100 permutations, qf(0.05, 1, 9)), `--workload c4` the whole-genome
configuration (24 chromosomes, 300 M kept bases, 100 samples, 1000 permutations) bin-packed over the
ranks (strong scaling, needs >= 2 GPUs for the coverage to be resident).  Multi-GPU (torchrun): every
rank owns a different synthetic chromosome of the same size (weak scaling, chromosome sharding); the
only exchange is the genome-wide pooling of mergeResults (R/mergeResults.R:216-235, 304-307,
380-386): an all-gather of the ranks' region areas and per-permutation maxima and an all-reduce of
the null histogram, once per step.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import math
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (chr_len, N kept bases, n samples, adjustment covariates, B permutations, cutoff p)
    "c2": dict(chr_len=48_129_895, N=4_800_000, n=12, extra=0, B=100, alpha=0.05,
               label="synthetic chr21 (48.1 Mbp, N=4.8M kept bases) 12 samples, 2-group model, 100 permutations"),
    "c3": dict(chr_len=249_250_621, N=25_000_000, n=60, extra=3, B=1000, alpha=1e-6,
               label="synthetic chr1 (249 Mbp, N=25M kept bases) 60 samples, 3 adjustment covariates, 1000 permutations"),
    "c3s": dict(chr_len=24_925_062, N=2_500_000, n=60, extra=3, B=1000, alpha=1e-6,
                label="1/10 of synthetic chr1 (N=2.5M kept bases) 60 samples, 3 adjustment covariates, 1000 permutations"),
    "tiny": dict(chr_len=2_000_000, N=200_000, n=12, extra=0, B=20, alpha=0.05,
                 label="tiny debug workload"),
    # BASELINE.json configs[3]: the whole genome after filtering, chromosome-sharded (strong scaling)
    "c4": dict(genome=True, N=300_000_000, n=100, extra=3, B=1000, alpha=1e-6,
               label="synthetic whole genome (24 chromosomes, 300M kept bases) 100 samples, 3 adjustment covariates, "
                     "1000 permutations, chromosomes bin-packed over the GPUs"),
    "c4s": dict(chr_len=10_000_000, N=1_000_000, n=100, extra=3, B=1000, alpha=1e-6,
                label="scaled slice of the whole-genome config (N=1M kept bases) 100 samples, 3 adjustment covariates, 1000 permutations"),
    "g_tiny": dict(genome=True, N=3_000_000, n=12, extra=0, B=20, alpha=0.05, label="tiny genome-mode debug workload"),
}

# hg19 chromosome lengths (tests/testthat/test_data.R:199 uses chr1's)
HG19 = {"chr1": 249250621, "chr2": 243199373, "chr3": 198022430, "chr4": 191154276, "chr5": 180915260,
        "chr6": 171115067, "chr7": 159138663, "chr8": 146364022, "chr9": 141213431, "chr10": 135534747,
        "chr11": 135006516, "chr12": 133851895, "chr13": 115169878, "chr14": 107349540, "chr15": 102531392,
        "chr16": 90354753, "chr17": 81195210, "chr18": 78077248, "chr19": 59128983, "chr20": 63025520,
        "chr21": 48129895, "chr22": 51304566, "chrX": 155270560, "chrY": 59373566}


def genome_chromosomes(cfg):
    """Per-chromosome configs of a genome workload: kept bases proportional to the chromosome length."""
    total = float(sum(HG19.values()))
    out = {}
    for name, length in HG19.items():
        c = dict(cfg)
        c.pop("genome")
        c["chr_len"] = length
        c["N"] = int(round(cfg["N"] * length / total / 64.0)) * 64
        out[name] = c
    return out


# ---------------------------------------------------------------------------------------------
# synthetic coverage (post-filter, like BASELINE.md's C4): expressed islands, piecewise-constant
# per-sample runs (read-length-like), Poisson counts, 5% of islands carry a 2x group effect


def island_layout(rng, chr_len, N):
    """island lengths (geometric, mean 200 bp) summing to N and the gaps between them."""
    lens = rng.geometric(1.0 / 200.0, size=int(N / 200 * 1.2) + 16)
    cs = np.cumsum(lens)
    k = int(np.searchsorted(cs, N))
    lens = lens[: k + 1].copy()
    lens[-1] -= cs[k] - N
    lens = lens[lens > 0]
    gaps = rng.geometric(len(lens) / max(chr_len - N, len(lens)), size=len(lens)) + 1
    scale = (chr_len - N) / gaps.sum()
    gaps = np.maximum(1, np.floor(gaps * scale)).astype(np.int64)
    gaps[-1] += chr_len - N - gaps.sum() if gaps.sum() < chr_len - N else 0
    return lens.astype(np.int64), gaps


def position_runs(lens, gaps, chr_len):
    """logical Rle: FALSE gap, TRUE island, ...; trailing FALSE to reach chr_len."""
    runs = np.empty(2 * len(lens), dtype=np.int64)
    runs[0::2] = gaps
    runs[1::2] = lens
    tail = chr_len - runs.sum()
    if tail > 0:
        runs = np.concatenate((runs, [tail]))
    else:
        runs[0] += tail  # shrink the first gap if rounding overshot
    return runs.astype(np.int32), 0


def make_workload(cfg, seed, torch=None, device=None):
    """Returns dict(cov [n x N] int32 (torch cuda tensor or numpy), pos_len, pos_first, models, group)."""
    rng = np.random.default_rng(seed)
    N, n = cfg["N"], cfg["n"]
    lens, gaps = island_layout(rng, cfg["chr_len"], N)
    pos_len, pos_first = position_runs(lens, gaps, cfg["chr_len"])
    nisl = len(lens)
    log_expr = rng.normal(3.0, 1.5, size=nisl).clip(-1, 8)
    effect = np.where(rng.random(nisl) < 0.05, 2.0, 1.0)
    depth = rng.uniform(0.6, 1.6, size=n)
    group = np.array([0] * (n // 2) + [1] * (n - n // 2))
    runlen = rng.integers(36, 77, size=n)
    offs = rng.integers(0, 76, size=n)
    starts = np.concatenate(([0], np.cumsum(lens)[:-1]))
    if torch is not None:
        dev = device
        isl = torch.repeat_interleave(torch.arange(nisl, device=dev), torch.as_tensor(lens, device=dev))
        lam = torch.as_tensor(np.exp(log_expr), device=dev, dtype=torch.float32)[isl]
        eff = torch.as_tensor(effect, device=dev, dtype=torch.float32)[isl]
        isl_start = torch.as_tensor(starts, device=dev)[isl]
        base = torch.arange(N, device=dev)
        g = torch.Generator(device=dev)
        g.manual_seed(seed)
        cov = torch.empty((n, N), dtype=torch.int32, device=dev)
        for s in range(n):
            # piecewise-constant runs (read-length-like); a run never crosses an island boundary: it
            # restarts at the island's first base
            first = ((base + int(offs[s])) // int(runlen[s])) * int(runlen[s]) - int(offs[s])
            first = torch.maximum(first, isl_start)
            rate = lam * float(depth[s]) * (eff if group[s] == 1 else 1.0)
            draw = torch.poisson(rate, generator=g)
            cov[s] = draw[first].to(torch.int32)
        depth_cov = torch.log2(cov.sum(dim=1).double() + 32.0).cpu().numpy()
    else:
        isl = np.repeat(np.arange(nisl), lens)
        lam, eff = np.exp(log_expr)[isl], effect[isl]
        isl_start = starts[isl]
        base = np.arange(N)
        cov = np.empty((n, N), dtype=np.int32)
        for s in range(n):
            first = np.maximum(((base + offs[s]) // runlen[s]) * runlen[s] - offs[s], isl_start)
            rate = lam * depth[s] * np.where(group[s] == 1, eff, 1.0)
            draw = rng.poisson(rate)
            cov[s] = draw[first]
        depth_cov = np.log2(cov.sum(axis=1).astype(np.float64) + 32.0)
    cols = [np.ones(n), group.astype(np.float64), depth_cov]
    cols += [rng.normal(size=n) for _ in range(cfg["extra"])]
    mod = np.column_stack(cols)
    mod0 = mod[:, [0] + list(range(2, mod.shape[1]))]
    return dict(cov=cov, pos_len=pos_len, pos_first=pos_first, mod=np.asfortranarray(mod),
                mod0=np.asfortranarray(mod0), group=group)


def theoretical_cutoff(alpha, mod, mod0):
    from scipy.stats import f as fdist
    n, p, p0 = mod.shape[0], mod.shape[1], mod0.shape[1]
    return float(fdist.isf(alpha, p - p0, n - p))


# ---------------------------------------------------------------------------------------------
# clocks sampler (NVML) -- runs DURING the timed region


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag = index, [], set(), False
        self.max_mhz = None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {
                getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
                getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
                getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
                getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
            }
            while not self.stop_flag:
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
                time.sleep(0.010)  # NVML queries take the driver lock: keep them sparse next to a latency-sensitive host loop
        except Exception as e:  # NVML missing: report that instead of inventing clocks
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def result(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ---------------------------------------------------------------------------------------------
# CPU baseline: the reference's algorithm shape on host cores (numpy/OpenBLAS port, "kind: port")


CPU_CHUNK = 100_000  # rows per chunk: the reference's documented chunksize for bplapply (users guide :1327)
_cpu_pool = None


def cpu_workers():
    """Worker count of the CPU arm: every host core the process may use, whatever the launcher set
    OMP_NUM_THREADS to (torchrun exports OMP_NUM_THREADS=1)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def _cpu_fstats_chunked(Y, mod, mod0, adjustF=0.0):
    """fstats.apply over 1e5-row chunks on a pool of host threads -- the shape of the reference's
    bplapply(chunks, fstats.apply, BPPARAM = SnowParam(mc.cores)) (R/calculateStats.R:105-127,
    R/calculatePvalues.R:325-329).  Each chunk is two dense residual-projector products (BLAS, one
    thread per worker: numpy releases the GIL inside them) + row sums."""
    global _cpu_pool
    from concurrent.futures import ThreadPoolExecutor
    from oracle import derfinder_oracle as O
    if _cpu_pool is None:
        _cpu_pool = ThreadPoolExecutor(max_workers=cpu_workers())
    P1, P0 = O._residual_projector(mod), O._residual_projector(mod0)
    n, df1, df0 = Y.shape[1], mod.shape[1], mod0.shape[1]

    def one(lo):
        y = Y[lo:lo + CPU_CHUNK]
        r0, r1 = y @ P0, y @ P1
        rss0 = np.einsum("ij,ij->i", r0, r0)
        rss1 = np.einsum("ij,ij->i", r1, r1)
        with np.errstate(divide="ignore", invalid="ignore"):
            return ((rss0 - rss1) / (df1 - df0)) / (adjustF + rss1 / (n - df1))

    parts = list(_cpu_pool.map(one, range(0, Y.shape[0], CPU_CHUNK)))
    return np.concatenate(parts) if parts else np.zeros(0)


def cpu_reference_pass(cov_nN, pos_len, pos_first, mod, mod0, perm, cutoff):
    """One pass of the same hot path the way the reference does it: log2 transform, per 1e5-row
    chunk dense residual-projector products Y %*% P1, Y %*% P0 (BLAS) on all host cores, rowSums, F;
    serial permutation loop; run segmentation + Rle view sums; .calcPval.  Returns #regions (sanity)."""
    from threadpoolctl import threadpool_limits
    from oracle import derfinder_oracle as O
    vals = (np.arange(len(pos_len)) % 2 == 0) == bool(pos_first)
    position = (vals, pos_len.astype(np.int64))
    with threadpool_limits(limits=1):  # one BLAS thread per chunk worker, the same at every --gpus N
        # preprocessCoverage: R's vectorised log2 and Reduce("+") (the oracle's exact-libm table look-up is a
        # test device, not how the reference spends its time)
        cov = np.ascontiguousarray(cov_nN.T)
        prep = dict(coverageProcessed=np.log2(cov + 32.0), mclapplyIndex=O.chunk_lengths(cov.shape[0], CPU_CHUNK, 1),
                    position=position, meanCoverage=O._ordered_row_sum(cov.T) / cov.shape[1], groupMeans={})
        F = _cpu_fstats_chunked(prep["coverageProcessed"], mod, mod0)
        res = O.calculate_pvalues(prep, mod, mod0, F, perm, cutoff, fstat_fn=_cpu_fstats_chunked)
    return 0 if res["regions"] is None else len(res["regions"]["start"])


def cpu_sample(work, cfg, perm, cutoff, n_bases, n_perm):
    cov = work["cov"]
    cov = cov[:, :n_bases].cpu().numpy() if hasattr(cov, "cpu") else cov[:, :n_bases]
    # position restricted to the first n_bases kept bases
    pl = work["pos_len"].astype(np.int64)
    vals = (np.arange(len(pl)) % 2 == 0) == bool(work["pos_first"])
    kept = np.cumsum(np.where(vals, pl, 0))
    k = int(np.searchsorted(kept, n_bases))
    pl2 = pl[: k + 1].copy()
    pl2[k] -= kept[k] - n_bases
    return cov, pl2.astype(np.int32), work["pos_first"], perm[:n_perm]


def run_reference_arm(args, cfg):
    """`--impl reference`: the reference's CPU implementation of the path (numpy port; R is not in
    this image) with all host threads OpenBLAS can use, on a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from derfinder_b200.rrng import permutation_indices
    if cfg.get("genome"):
        cfg = dict(genome_chromosomes(cfg)["chr1"])
    nb, npm = min(cfg["N"], args.cpu_bases), min(cfg["B"], args.cpu_perms)
    # the bounded sample: the first `nb` kept bases of a chromosome of the workload's shape (same generator, same
    # island statistics; generating all 25 M bases x 60 samples on the host would take longer than the arm itself)
    scale = nb / float(cfg["N"])
    cfg = dict(cfg, N=nb, chr_len=max(int(cfg["chr_len"] * scale), nb + 1000))
    work = make_workload(cfg, 20140923 + 2)
    perm_all = permutation_indices(range(1, cfg["B"] + 1), cfg["n"], "Rejection")
    cutoff = theoretical_cutoff(cfg["alpha"], work["mod"], work["mod0"])
    cov, pl, pf, perm = cpu_sample(work, cfg, perm_all, cutoff, nb, npm)
    for _ in range(args.warmup):
        cpu_reference_pass(cov[:, : nb // 10], *cpu_sample(work, cfg, perm_all, cutoff, nb // 10, 2)[1:3], work["mod"],
                           work["mod0"], perm[:2], cutoff)
    steps = max(1, min(args.steps, 10))  # every step is ~4 s of CPU work: keep the arm within a few minutes
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_reference_pass(cov, pl, pf, work["mod"], work["mod0"], perm, cutoff)
    dt = (time.perf_counter() - t0) / steps
    value = nb * (npm + 1) / dt
    cores = cpu_workers()
    sample = (f"first {nb} kept bases x {npm} permutations of the workload per step (numpy/OpenBLAS port, 1e5-row "
              f"chunks on {cores} host threads)")
    line = {"impl": "reference", "metric": "base_x_permutation_fstats_per_sec", "value": value,
            "unit": "base*perm F-stats/s", "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
            "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": {"workload": WORKLOADS[args.workload]["label"], "sample": sample},
            "cpu_baseline": {"value": value, "unit": "base*perm F-stats/s", "cores": cores, "kind": "port",
                             "sample": sample},
            "e2e": {"value": value, "unit": "base*perm F-stats/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
# our arm


class Step:
    """The hot path through the C ABI (ctypes).  Device-resident and host-buffer variants."""

    def __init__(self, eng, work, perm, cutoff, torch):
        from derfinder_b200 import _lib as L
        self.L, self.eng, self.lib, self.torch = L, eng, eng.lib, torch
        self.work, self.perm, self.cut = work, np.ascontiguousarray(perm, dtype=np.int32), cutoff
        self.n, self.N = work["cov"].shape
        self.mod, self.mod0 = work["mod"], work["mod0"]
        self.pl = np.ascontiguousarray(work["pos_len"], dtype=np.int32)
        self._F_host = np.empty(self.N)  # the caller's fstats vector (R: a numeric vector it already owns)
        self.rle = None
        self.sync_f = False  # True: dfb_fstats blocks until the F-statistics are in host memory
        self.last_breakdown = {}

    def _fp(self, a):
        return a.ctypes.data_as(self.L.c_f64p)

    def set_rle_input(self, host_cov):
        """Run-length encode the [n x N] host coverage once (outside any timed region): this is the
        form in which the reference holds coverage (DataFrame of Rle, R/loadCoverage.R:323-326)."""
        vals, lens = [], []
        for s in range(self.n):
            x = host_cov[s]
            brk = np.flatnonzero(x[1:] != x[:-1]) + 1
            starts = np.concatenate(([0], brk))
            vals.append(np.ascontiguousarray(x[starts], dtype=np.int32))
            lens.append(np.diff(np.concatenate((starts, [x.size]))).astype(np.int32))
        # the step's inputs sit in page-locked host memory (the bench contract): the library DMAs such columns
        # where they lie; pageable columns (an R session's vectors) go through its pinned staging ring instead
        # (--e2e-pageable measures that path)
        if not getattr(self, "pageable", False):
            keep = []
            for lst in (vals, lens):
                for i, a in enumerate(lst):
                    t = self.torch.empty(a.shape, dtype=self.torch.int32, pin_memory=True)
                    v = t.numpy()
                    v[...] = a
                    keep.append(t)
                    lst[i] = v
            self._pinned_keep = keep
        vp = (C.c_void_p * self.n)(*[v.ctypes.data for v in vals])
        lp = (C.c_void_p * self.n)(*[l.ctypes.data for l in lens])
        nr = np.asarray([len(v) for v in vals], dtype=np.int64)
        self.rle = (vals, lens, vp, lp, nr)

    def device_pass(self, keep=False):
        L, lib, eng = self.L, self.lib, self.eng
        cov = C.c_void_p()
        # the coverage is resident in HBM in the engine's layout: the handle reads it in place
        L.check(lib.dfb_cov_view_dense_dev(eng.h, self.n, self.N, L.DFB_I32, C.c_void_p(self.work["cov"].data_ptr()),
                                           self.N, L.ptr(self.pl, C.c_int32), len(self.pl), self.work["pos_first"],
                                           C.byref(cov)))
        try:
            L.check(lib.dfb_preprocess(eng.h, cov, 32.0, None, 0))
            L.check(lib.dfb_fstats(eng.h, cov, self._fp(self.mod), self.mod.shape[1], self._fp(self.mod0),
                                   self.mod0.shape[1], 0.0, None))
            rh, nh = C.c_void_p(), C.c_void_p()
            rc = L.check(lib.dfb_calculate_pvalues(eng.h, cov, None, self._fp(self.mod), self.mod.shape[1],
                                                   self._fp(self.mod0), self.mod0.shape[1], 0.0,
                                                   L.ptr(self.perm, C.c_int32), self.perm.shape[0], -self.cut, self.cut,
                                                   0, 3000, C.byref(rh), C.byref(nh), None))
            out = (int(lib.dfb_regions_count(rh)) if rc == 0 else 0, int(lib.dfb_nulls_count(nh)) if rc == 0 else 0)
            if keep:  # no observed regions: the chromosome takes part in the pooling with NULL handles
                return (out, rh, nh, cov) if rc == 0 else (out, None, None, cov)
            if rc == 0:
                lib.dfb_regions_free(rh)
                lib.dfb_nulls_free(nh)
            return out
        finally:
            if not keep:
                lib.dfb_cov_free(cov)

    def host_pass(self, host_cov):
        """Reference-facing calls with host buffers: upload coverage, download everything a
        calculatePvalues() caller receives."""
        L, lib, eng = self.L, self.lib, self.eng
        cov = C.c_void_p()
        tm = {}
        t0 = time.perf_counter()

        def lap(name):
            nonlocal t0
            t1 = time.perf_counter()
            tm[name] = tm.get(name, 0.0) + (t1 - t0) * 1e3
            t0 = t1

        if self.rle is not None:
            # the reference's own input type: a DataFrame of integer Rle columns (values + run lengths)
            vals, lens, vp, lp, nr = self.rle
            L.check(lib.dfb_cov_from_rle(eng.h, self.n, L.DFB_I32, vp, lp, L.ptr(nr, C.c_int64), self.N,
                                         L.ptr(self.pl, C.c_int32), len(self.pl), self.work["pos_first"],
                                         C.byref(cov)))
            cov_bytes = sum(v.nbytes + l.nbytes for v, l in zip(vals, lens))
        else:
            L.check(lib.dfb_cov_from_dense(eng.h, self.n, self.N, L.DFB_I32, host_cov.ctypes.data_as(C.c_void_p),
                                           self.N, L.ptr(self.pl, C.c_int32), len(self.pl), self.work["pos_first"],
                                           C.byref(cov)))
            cov_bytes = host_cov.nbytes
        lap("upload_coverage")
        d2h = 0
        try:
            L.check(lib.dfb_preprocess(eng.h, cov, 32.0, None, 0))
            lap("preprocess")
            F = self._F_host
            # the F-statistics are downloaded into the caller's vector in the background while
            # calculatePvalues (which uses their device copy) runs; the wait below is inside the timed region
            fstats_call = lib.dfb_fstats if self.sync_f else lib.dfb_fstats_begin
            L.check(fstats_call(eng.h, cov, self._fp(self.mod), self.mod.shape[1], self._fp(self.mod0),
                                self.mod0.shape[1], 0.0, self._fp(F)))
            d2h += F.nbytes
            lap("calculateStats+download_F" if self.sync_f else "calculateStats(F download started)")
            rh, nh = C.c_void_p(), C.c_void_p()
            # `fstats` is the object calculateStats() just returned: its device copy is still cached in
            # the coverage handle, so the glue passes NULL instead of uploading it again
            rc = L.check(lib.dfb_calculate_pvalues(eng.h, cov, None, self._fp(self.mod), self.mod.shape[1],
                                                   self._fp(self.mod0), self.mod0.shape[1], 0.0,
                                                   L.ptr(self.perm, C.c_int32), self.perm.shape[0], -self.cut, self.cut,
                                                   0, 3000, C.byref(rh), C.byref(nh), None))
            lap("calculatePvalues")
            if not self.sync_f:
                L.check(lib.dfb_fstats_wait(eng.h))
                lap("wait_for_F_download")
            if rc == 0:
                R, M = int(lib.dfb_regions_count(rh)), int(lib.dfb_nulls_count(nh))
                st, en, ist, ien, cl = (np.empty(R, dtype=np.int32) for _ in range(5))
                val, area, cll, pv, mc = (np.empty(R) for _ in range(5))
                L.check(lib.dfb_regions_get(rh, L.ptr(st, C.c_int32), L.ptr(en, C.c_int32), self._fp(val),
                                            self._fp(area), L.ptr(ist, C.c_int32), L.ptr(ien, C.c_int32),
                                            L.ptr(cl, C.c_int32), self._fp(cll)))
                L.check(lib.dfb_regions_get_pvalues(rh, self._fp(pv)))
                L.check(lib.dfb_regions_view_means(eng.h, rh, cov, -1, self._fp(mc)))
                # caller-owned result vectors, reused across steps (grown when needed)
                if getattr(self, "_null_cap", -1) < 2 * M:
                    self._null_cap = int(2 * M * 1.1) + 16
                    self._nbuf = (np.empty(self._null_cap), np.empty(self._null_cap),
                                  np.empty(self._null_cap, dtype=np.int32), np.empty(self._null_cap, dtype=np.int32))
                ns, na, nw, npid = (a[:2 * M] for a in self._nbuf)
                L.check(lib.dfb_nulls_get(nh, 1, self._fp(ns), L.ptr(nw, C.c_int32), self._fp(na),
                                          L.ptr(npid, C.c_int32)))
                d2h += R * (5 * 4 + 5 * 8) + M * (8 + 8 + 4 + 4)
                lib.dfb_regions_free(rh)
                lib.dfb_nulls_free(nh)
                lap("download_regions_nulls")
            h2d = cov_bytes + self.pl.nbytes + self.perm.nbytes
            self.last_breakdown = tm
            return h2d, d2h
        finally:
            lib.dfb_cov_free(cov)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-bases", type=int, default=1_600_000, help="bounded CPU sample: kept bases per step")
    ap.add_argument("--cpu-perms", type=int, default=20, help="bounded CPU sample: permutations per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="developer: skip the end-to-end pass (the line has no e2e key)")
    ap.add_argument("--e2e-input", default="rle", choices=["rle", "dense"],
                    help="host coverage format of the end-to-end pass: Rle columns (the reference's type) or dense")
    ap.add_argument("--e2e-pageable", action="store_true",
                    help="end-to-end pass: Rle columns in pageable host memory (the library's pinned staging ring) instead of "
                         "page-locked memory (direct DMA)")
    ap.add_argument("--e2e-sync-f", action="store_true",
                    help="end-to-end pass: blocking dfb_fstats download instead of dfb_fstats_begin / dfb_fstats_wait")
    ap.add_argument("--opt", action="append", default=[], help="engine option key=value (developer knob)")
    ap.add_argument("--as-rank", type=int, default=None,
                    help="developer: generate the synthetic chromosome rank R would own (single-GPU runs)")
    ap.add_argument("--force-pool", action="store_true",
                    help="developer: run the genome-wide pooling step on one GPU too (its local cost without NCCL)")
    ap.add_argument("--screen", type=int, default=None, help="override the engine's tcgen05 screening option")
    args = ap.parse_args()
    cfg = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference_arm(args, cfg)
        return

    import torch
    import torch.distributed as dist

    import derfinder_b200 as dfb
    from derfinder_b200 import _lib as L
    from derfinder_b200 import pool
    from derfinder_b200.rrng import permutation_indices

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    W = max(args.warmup, 3)

    eng = dfb.Engine(local, stream=torch.cuda.current_stream().cuda_stream)
    if args.screen is not None:
        eng.set_option("screen", args.screen)
    for kv in args.opt:
        key, val = kv.split("=")
        eng.set_option(key, float(val))
    if world > 1:
        pool.init_communicator(eng, dist, device)  # the library's own NCCL communicator (id via the process group)
    genome = bool(cfg.get("genome"))
    n, B = cfg["n"], cfg["B"]
    perm = permutation_indices(range(1, B + 1), n, "Rejection")  # same relabellings on every rank
    if genome:
        # strong scaling: the 24 chromosomes of one genome, longest first onto the least loaded rank
        chr_cfg = genome_chromosomes(cfg)
        mine = pool.assign_chromosomes({k: v["N"] for k, v in chr_cfg.items()}, world)[rank]
        need_gb = sum(chr_cfg[k]["N"] for k in mine) * n * 4e-9 + max(chr_cfg[k]["N"] for k in chr_cfg) * n * 16e-9
        if need_gb > 150:
            raise SystemExit("workload %s needs %.0f GB on this rank: run it on more GPUs" % (args.workload, need_gb))
        steps_obj = []
        for k in mine:
            w = make_workload(chr_cfg[k], 20140923 + 4 + 7919 * list(HG19).index(k), torch=torch, device=device)
            steps_obj.append((k, w))
        # one design for the whole genome (the samples are the same on every chromosome)
        gm = make_workload(dict(chr_cfg["chr21"], N=64000, chr_len=640000), 20140923 + 4)
        cut = theoretical_cutoff(cfg["alpha"], gm["mod"], gm["mod0"])
        chroms = []
        for k, w in steps_obj:
            w["mod"], w["mod0"] = gm["mod"], gm["mod0"]
            chroms.append(Step(eng, w, perm, cut, torch))
        N = sum(chr_cfg[k]["N"] for k in chr_cfg)  # units of the whole job
        N_mine = sum(chr_cfg[k]["N"] for k in mine)
        work = gm
    else:
        # every rank owns its own synthetic chromosome (same shape): weak scaling
        data_rank = rank if args.as_rank is None else args.as_rank
        work = make_workload(cfg, 20140923 + 2 + 1000 * data_rank, torch=torch, device=device)
        N = cfg["N"]
        N_mine = N
        cut = theoretical_cutoff(cfg["alpha"], work["mod"], work["mod0"])
        chroms = [Step(eng, work, perm, cut, torch)]
        mine = ["chr"]
    step = chroms[0]
    step.sync_f = args.e2e_sync_f
    step.pageable = args.e2e_pageable
    pooled = world > 1 or args.force_pool or genome

    timed_pass = None  # per-step events around this rank's own chromosome passes (timed region only)

    def one_step():
        if not pooled:
            return step.device_pass()
        if timed_pass is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        kept, Rs, Ms = [], 0, 0
        for st in chroms:
            (R_c, M_c), rh, nh, cov = st.device_pass(keep=True)
            kept.append((rh, nh, cov))
            Rs += R_c
            Ms += M_c
        if timed_pass is not None:
            e1.record()
            timed_pass.append((e0, e1))
        # mergeResults over all GPUs: ONE library call, once per rank after its last chromosome
        pool.merge_results(eng, [(rh, nh) for rh, nh, _ in kept], B, cut, merge_area=True)
        for rh, nh, cov in kept:
            if rh is not None:
                eng.lib.dfb_regions_free(rh)
                eng.lib.dfb_nulls_free(nh)
            eng.lib.dfb_cov_free(cov)
        return Rs, Ms

    for _ in range(W):
        R, M = one_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    eng.reset_counters()
    # live CUDA-event brackets around the dominant kernel only (every K2 launch of the timed region); the
    # other families are timed in a separate short pass below so that their ~50 event pairs per step
    # do not sit inside the measurement
    if not os.environ.get("BENCH_NO_KTIMER"):
        eng.enable_timing(families=[L.K_FSTAT_TC])
    sampler = ClockSampler(local)
    if not os.environ.get("BENCH_NO_SAMPLER"):
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record()
    timed_pass = []
    step_wall = []
    for _ in range(args.steps):
        t_s = time.perf_counter()
        R, M = one_step()
        step_wall.append((time.perf_counter() - t_s) * 1e3)
    ev1.record()
    torch.cuda.synchronize()
    own_ms = sum(a.elapsed_time(b) for a, b in timed_pass) / max(len(timed_pass), 1)
    timed_pass = None
    sampler.stop_flag = True
    if sampler.is_alive():
        sampler.join()
    ms = ev0.elapsed_time(ev1) / args.steps
    if os.environ.get("DFB_GAPS"):
        print("[bench] host wall per timed step (ms):", [round(x, 2) for x in step_wall], file=sys.stderr)
    launches = eng.launch_count()
    kt = {L.K_FSTAT_TC: eng.kernel_time(L.K_FSTAT_TC)}
    eng.enable_timing(False)
    # breakdown pass (not part of `value`): every family bracketed, a few steps
    prof_steps = max(2, min(10, args.steps)) if not genome else 1
    eng.reset_counters()
    eng.enable_timing(True)
    for _ in range(prof_steps):
        one_step()
    torch.cuda.synchronize()
    for k in (L.K_PREP, L.K_FSTAT, L.K_REGIONS, L.K_PVAL):
        t_ms, t_l = eng.kernel_time(k)
        kt[k] = (t_ms * args.steps / prof_steps, t_l * args.steps / prof_steps)
    eng.enable_timing(False)
    own_all = [own_ms]
    if world > 1:
        t = torch.tensor([ms], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        own = torch.zeros(world, device=device, dtype=torch.float64)
        own[rank] = own_ms
        dist.all_reduce(own, op=dist.ReduceOp.SUM)
        own_all = [round(v, 4) for v in own.tolist()]
        dist.barrier()

    # end to end through host buffers (pinned), same pass
    h2d = d2h = 0
    e2e_s = float("nan")
    if genome:
        args.no_e2e = True  # the genome workload is the strong-scaling arm; the host-buffer pass is measured on c3
    if not args.no_e2e:
        host_cov = torch.empty((n, N), dtype=torch.int32, pin_memory=True)
        host_cov.copy_(work["cov"])
        host_np = host_cov.numpy()
        if args.e2e_input == "rle":
            step.set_rle_input(host_np)
        step.host_pass(host_np)
        torch.cuda.synchronize()
        ke = max(2, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(ke):
            h2d, d2h = step.host_pass(host_np)
        torch.cuda.synchronize()
        e2e_s = (time.perf_counter() - t0) / ke
        if os.environ.get("DFB_GAPS"):  # developer: per-launch trace of one end-to-end pass
            print("[bench] end-to-end pass trace", file=sys.stderr)
            eng.enable_timing(True)
            step.host_pass(host_np)
            eng.enable_timing(False)
            print("[bench] end-to-end breakdown (ms):", step.last_breakdown, file=sys.stderr)
    if world > 1:
        t = torch.tensor([e2e_s], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())

    units = float(N) * (B + 1) * (1 if genome else world)  # a genome is one fixed job; else one chromosome per rank
    value = units / (ms * 1e-3)
    line = {"metric": "base_x_permutation_fstats_per_sec", "value": value, "unit": "base*perm F-stats/s",
            "n_gpus": world, "steps": args.steps, "warmup": W, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong" if genome else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": cfg["label"], "N": N, "n": n, "p": int(work["mod"].shape[1]),
                       "p0": int(work["mod0"].shape[1]), "B": B, "f_cutoff": cut,
                       "sharding": ("24 chromosomes bin-packed over the GPUs (longest first), mergeResults pooling once "
                                    "per step" if genome else "one chromosome per GPU, mergeResults pooling once per step"
                                    if world > 1 else "single GPU"),
                       "l2": "inputs larger than L2 (coverage %.0f MB + log2 matrix %.0f MB per rank and step; no flush)"
                             % (4e-6 * n * N_mine, 8e-6 * n * N_mine),
                       "regions_rank0": R, "null_regions_rank0": M},
            "gpu_launches": int(launches),
            "e2e": {"value": units / e2e_s, "unit": "base*perm F-stats/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_s * 1e3,
                    "last_step_breakdown_ms": {k: round(v, 3) for k, v in step.last_breakdown.items()}},
            "clocks": sampler.result()}
    if args.no_e2e:
        del line["e2e"]
    if world > 1:
        # every rank's own chromosome pass (different synthetic chromosomes: the slowest bounds the step)
        # and what the genome-wide pooling (two all-gathers + ranking) adds on top of it
        line["multi_gpu"] = {"chromosome_pass_ms_per_rank": own_all,
                             "pooling_ms": round(ms - max(own_all), 4),
                             "collectives": "dfb_merge_results on the library's NCCL communicator: all-gather of "
                                            "{region and null counts}, all-gather of the records {per-permutation max "
                                            "null area, this rank's region areas}, all-reduce(sum) of the null "
                                            "histogram over the merged region list; the null lists stay on their GPU"}
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        # dominant kernel: the permutation F-stat scan (fstat_null_kernel, tcgen05 screen + FP64
        # refinement; its B-operand builder is counted with it).  Algorithmic work 2*N*n*p
        # FP64-equivalent flops per (base, labelling) over the B permutations of the step.
        p = int(work["mod"].shape[1])
        fst_ms, fst_l = kt[L.K_FSTAT]
        tc_ms, tc_l = kt[L.K_FSTAT_TC]
        k2_ms = tc_ms / args.steps
        flops = 2.0 * N_mine * n * p * B  # what rank 0's own K2 launches (the timed ones) computed
        achieved = flops / (k2_ms * 1e-3) / 1e12 if k2_ms > 0 else None
        # K2 is timed inside a long step (tens of 20 ms steps back to back, SM clocks under the power cap): the
        # sustained bf16 figure of MEASURED_PEAKS.json is its roofline; the burst figure is kept beside it
        peak_burst = peaks.get("bf16_tflops", 1590.0)
        peak = peaks.get("bf16_tflops_sustained", peak_burst)
        hbm = peaks.get("hbm_gbs", 6650.0)
        alg_bytes = 8.0 * N_mine * n  # read Y once (SURVEY 8d lower bound for the null scan)
        # DRAM traffic per step of the K2 launches: from the committed summary of an `ncu --set full` capture of
        # this same command (profiles/k2_traffic.json, written by tools/ncu_summary.py), else null
        traffic, traffic_src = None, None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "k2_traffic.json")))
            if args.workload in tj:
                traffic = float(tj[args.workload]["dram_bytes_per_step"])
                traffic_src = tj[args.workload].get("source")
        except Exception:
            pass
        if n <= 16:
            why = ("n=%d, p=%d: %d flop per (base, permutation) and K = 16 of a 64-wide operand row; the kernel is "
                   "bound by the TMEM epilogue and the FP64 refinement of the exceedances (a few percent of all "
                   "pairs at this cutoff), not by the tensor pipe or HBM (DESIGN.md section 4)" % (n, p, 2 * n * p))
        else:
            why = ("n=%d, p=%d: one fp16 tcgen05 product per (128-base tile, 32-permutation chunk) with the constant "
                   "basis column dropped -- executed flops are %.2fx the algorithmic 2*N*n*p*B; the chunk loop is "
                   "bound by the MMA -> TMEM scan -> MMA round trip of an accumulator stage (tensor pipe ~40-50%% "
                   "active in the ncu capture), see DESIGN.md section 4" %
                   (n, p, (((n + 15) // 16) * 16) * (p - 1) / float(n * p)))
        line["roofline"] = {"bound": "tensor", "kernel": "fstat_null2_kernel (K2 permutation scan)", "achieved": achieved,
                            "peak": peak, "unit": "TFLOP/s", "frac": (achieved / peak if achieved else None),
                            "traffic": traffic,
                            "peak_source": ("MEASURED_PEAKS.json bf16 sustained (kernel timed inside a long step)"
                                            if "bf16_tflops_sustained" in peaks else
                                            "MEASURED_PEAKS.json bf16 burst" if "bf16_tflops" in peaks else "fallback"),
                            "frac_of_burst_peak": (achieved / peak_burst if achieved else None), "burst_peak": peak_burst,
                            "algorithmic_flops_per_step": flops, "kernel_ms_per_step": k2_ms,
                            "launches_per_step": tc_l / args.steps,
                            "kernel_share_of_step": k2_ms / ms,
                            "traffic_source": traffic_src,
                            "notes": {"what_bounds_it": why,
                                      "k2_family": "K2 time = the null-scan launches plus their operand builders "
                                                   "(permuted basis tiles, per-base thresholds) measured by CUDA "
                                                   "events on the engine's stream inside the timed region",
                                      "hbm_equivalent_GBps": alg_bytes / (k2_ms * 1e-3) / 1e9 if k2_ms > 0 else None,
                                      "hbm_equivalent_frac": (alg_bytes / (k2_ms * 1e-3) / 1e9 / hbm) if k2_ms > 0 else None,
                                      "base_x_perm_per_s_kernel_only": N_mine * float(B) / (k2_ms * 1e-3) if k2_ms > 0 else None,
                                      "other_kernels": "timed in a separate pass of %d steps with every launch "
                                                       "bracketed (not inside the timed region); " % prof_steps +
                                                       "regions / pval include the observed-region chain, which runs on "
                                                       "an auxiliary stream concurrently with the null scan: its event "
                                                       "brackets contain time spent queued behind that kernel, so these "
                                                       "two sums overlap the K2 time instead of adding to it"},
                            "other_kernels_ms_per_step": {"observed_F": fst_ms / args.steps,
                                                          "prep": kt[L.K_PREP][0] / args.steps,
                                                          "regions": kt[L.K_REGIONS][0] / args.steps,
                                                          "pval": kt[L.K_PVAL][0] / args.steps}}
        if world == 1 and not args.no_cpu_baseline and not genome:
            from oracle import c_oracle
            c_oracle.build()
            nb, npm = min(N, args.cpu_bases), min(B, args.cpu_perms)
            cov_s, pl_s, pf_s, perm_s = cpu_sample(work, cfg, perm, cut, nb, npm)
            cpu_reference_pass(cov_s[:, : nb // 10], *cpu_sample(work, cfg, perm, cut, nb // 10, 2)[1:3], work["mod"],
                               work["mod0"], perm_s[:2], cut)
            t0 = time.perf_counter()
            cpu_reference_pass(cov_s, pl_s, pf_s, work["mod"], work["mod0"], perm_s, cut)
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": nb * (npm + 1) / dt, "unit": "base*perm F-stats/s",
                                    "cores": cpu_workers(), "kind": "port", "seconds": dt,
                                    "sample": f"first {nb} kept bases x {npm} permutations of the same workload, "
                                              "numpy/OpenBLAS port of the reference's algorithm (1e5-row chunks "
                                              f"on {cpu_workers()} host threads)"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    eng.close()


if __name__ == "__main__":
    main()
