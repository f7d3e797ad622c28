"""CPU restatement (numpy) of derfinder's single-base F-statistic / expressed-region hot path.

TEST INFRASTRUCTURE ONLY.  This module is the *oracle* the CUDA engine is checked against; it is
imported by `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs and by nothing
in `derfinder_b200/`.  Every function cites the reference lines it restates (paths relative to the
upstream `lcolladotor/derfinder` checkout, v1.35.0).  Parity status: PINNED — see
`tests/test_oracle_golden.py`, which replays the reference's own golden objects
(`data/genomeFstats.RData`, `data/genomeRegions.RData`, `inst/extdata/chr21/*.Rdata` and the
literal known answers in `tests/testthat/*.R`) from the fixtures in `tests/golden/`.

Un-vendored dependencies whose published behaviour is restated here:
  * derfinderHelper::fstats.apply (Bioconductor, `DESCRIPTION:22` ">= 1.1.0", no lockfile):
    restated from the reference's own normative formula in `tests/testthat/test_Fstats.R:52-63`.
  * IRanges/S4Vectors `slice`, `reduce`, `disjoin`, `Views` sum/mean, `Rle` (`DESCRIPTION:27,34`).
  * base R `set.seed`/`sample` -> `oracle/rrng.py`.

Conventions: indices handed back to callers are 1-based where the reference's are (indexStart,
indexEnd, start, end, cluster ids, permutation ids) so that results can be compared with the
goldens without translation.  Coverage matrices are `[N x n]` (bases x samples) like R's DataFrame.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

# ---------------------------------------------------------------------------------------------
# Rle helpers (S4Vectors::Rle semantics)


def rle_encode(x):
    """values, lengths of the run-length encoding of a dense vector (adjacent equal values merge;
    NaN never equals NaN, -0.0 == 0.0 keeps the first representative, as in S4Vectors)."""
    x = np.asarray(x)
    if x.size == 0:
        return x[:0].copy(), np.zeros(0, dtype=np.int64)
    brk = np.flatnonzero(x[1:] != x[:-1]) + 1
    starts = np.concatenate(([0], brk))
    lens = np.diff(np.concatenate((starts, [x.size])))
    return x[starts].copy(), lens.astype(np.int64)


def rle_normalize(values, lengths):
    """Merge adjacent equal runs and drop zero-length runs (what the `Rle()` constructor does)."""
    values = np.asarray(values)
    lengths = np.asarray(lengths, dtype=np.int64)
    nz = lengths > 0
    values, lengths = values[nz], lengths[nz]
    if values.size == 0:
        return values, lengths
    new = np.concatenate(([True], values[1:] != values[:-1]))
    grp = np.cumsum(new) - 1
    out_len = np.zeros(int(grp[-1]) + 1, dtype=np.int64)
    np.add.at(out_len, grp, lengths)
    return values[new], out_len


def rle_decode(values, lengths):
    return np.repeat(np.asarray(values), np.asarray(lengths, dtype=np.int64))


def joint_runs(cols):
    """Common refinement of several Rle columns of equal total length.

    cols: list of (values, lengths).  Returns (vals [n x S], lens [S]) where every sample is
    constant on each of the S segments.  This is only a convenience for the oracle: R performs the
    same arithmetic run-by-run inside S4Vectors' Rle group generics."""
    ends = [np.cumsum(np.asarray(l, dtype=np.int64)) for _, l in cols]
    total = int(ends[0][-1]) if len(ends[0]) else 0
    for e in ends:
        if (int(e[-1]) if len(e) else 0) != total:
            raise ValueError("Rle columns differ in length")
    allends = np.unique(np.concatenate(ends)) if total else np.zeros(0, dtype=np.int64)
    lens = np.diff(np.concatenate(([0], allends)))
    vals = []
    for (v, _), e in zip(cols, ends):
        idx = np.searchsorted(e, allends, side="left")
        vals.append(np.asarray(v)[idx])
    return (np.stack(vals) if vals else np.zeros((0, 0))), lens


# ---------------------------------------------------------------------------------------------
# filterData  (R/filterData.R:89-241)


@dataclass
class FilterResult:
    coverage: np.ndarray | None  # [N x n] dense, filtered (normalised when totalMapped given)
    position: tuple | None  # (logical values, lengths) Rle over the input length, or None
    meanCoverage: np.ndarray | None  # [N]


def _ordered_row_sum(vals):
    """((x1 + x2) + x3) + ...  in sample order -- `Reduce("+", data)` (R/filterData.R:147,171)."""
    acc = vals[0].astype(np.float64).copy()
    for s in range(1, vals.shape[0]):
        acc = acc + vals[s]
    return acc


def filter_data(cols, cutoff=None, index=None, filter="one", totalMapped=None, targetSize=80e6,
                returnMean=False, returnCoverage=True):
    """filterData() on Rle columns.  cols: list of n (values, lengths) pairs, all of length
    `sum(index)` when a previous logical `index` Rle (values, lengths) is given, else of the
    chromosome length.  (R/filterData.R:119-238)"""
    if filter not in ("one", "mean"):
        raise ValueError("filter %in% c('one', 'mean') is not TRUE")
    vals, lens = joint_runs(cols)
    n = vals.shape[0]
    if totalMapped is not None and targetSize != 0:  # :119-130
        mapped = np.broadcast_to(np.asarray(totalMapped, dtype=np.float64) / targetSize, (n,))
        vals = vals.astype(np.float64) / mapped[:, None]
    mean_seg = None
    newmask = None
    if cutoff is not None:
        if filter == "one":  # :138-145
            newmask = np.zeros(vals.shape[1], dtype=bool)
            for s in range(n):
                newmask |= vals[s] > cutoff
        else:  # :146-156
            mean_seg = _ordered_row_sum(vals) / n
            newmask = mean_seg > cutoff
    # final index (:159-164)
    if cutoff is None:
        position = index
    elif index is not None:
        iv, il = np.asarray(index[0]).astype(bool), np.asarray(index[1], dtype=np.int64)
        dense_new = np.repeat(newmask, lens)
        # finalidx[index] <- newindex : walk the TRUE runs of the old index
        out_v, out_l = [], []
        off = 0
        for v, l in zip(iv, il):
            if v:
                sv, sl = rle_encode(dense_new[off:off + l])
                out_v.append(sv)
                out_l.append(sl)
                off += l
            else:
                out_v.append(np.array([False]))
                out_l.append(np.array([l], dtype=np.int64))
        position = rle_normalize(np.concatenate(out_v), np.concatenate(out_l))
    else:
        position = rle_normalize(newmask, lens)
    mean_out = None
    if returnMean:  # :168-178
        if mean_seg is None:
            mean_seg = _ordered_row_sum(vals) / n
        keep = newmask if newmask is not None else np.ones(vals.shape[1], dtype=bool)
        mean_out = np.repeat(mean_seg[keep], lens[keep])
    cov = None
    if returnCoverage:  # :180-197
        keep = newmask if newmask is not None else np.ones(vals.shape[1], dtype=bool)
        cov = np.repeat(vals[:, keep], lens[keep], axis=1).T.copy()
    return FilterResult(cov, position, mean_out)


# ---------------------------------------------------------------------------------------------
# preprocessCoverage  (R/preprocessCoverage.R:97-260)


def chunk_lengths(numrow: int, chunksize, mc_cores: int = 1):
    """split.len of R/preprocessCoverage.R:171-188, 218-227."""
    if chunksize is None:
        chunksize = math.ceil(numrow / mc_cores)
    lastloop = int(numrow // chunksize) if chunksize else 0
    if chunksize and numrow % chunksize == 0 and lastloop > 0:
        lastloop -= 1
    if lastloop == 0:
        return [numrow]
    out = [int(chunksize)] * lastloop
    rest = numrow - sum(out)
    if rest > 0:
        out.append(int(rest))
    return out


def log2_libm(x):
    """log2 through the C library (what R's log2() calls), not numpy's SIMD kernel, which can differ
    from glibc in the last bit."""
    x = np.asarray(x, dtype=np.float64)
    u, inv = np.unique(x, return_inverse=True)
    lut = np.array([math.log2(v) if v > 0 else (-np.inf if v == 0 else np.nan) for v in u.tolist()])
    return lut[inv].reshape(x.shape)


def preprocess_coverage(coverage, position, groupInfo=None, scalefac=32, chunksize=5e6,
                        meanCoverage=None, mc_cores=1):
    """coverage [N x n] (already filtered), position Rle.  Returns dict with the 5 reference names.
    `mclapplyIndex` is returned as the list of chunk lengths (chunk k is rows
    [sum(len[:k]), sum(len[:k+1])) ), from which the logical Rles of :248-250 follow."""
    cov = np.asarray(coverage)
    N, n = cov.shape
    if position is None:
        raise ValueError("'coverageInfo$position' is NULL when it shouldn't be.")
    means = meanCoverage
    if means is None:  # :191-193
        means = _ordered_row_sum(cov.T) / n
    groupMeans = {}
    if groupInfo is not None:  # :197-206  (levels in sorted order, like factor())
        g = np.asarray(groupInfo)
        for lev in sorted(set(g.tolist())):
            sel = np.flatnonzero(g == lev)
            groupMeans[lev] = _ordered_row_sum(cov[:, sel].T) / len(sel)
    Y = log2_libm(cov.astype(np.float64) + scalefac)  # :209-212
    return dict(coverageProcessed=Y, mclapplyIndex=chunk_lengths(N, chunksize, mc_cores),
                position=position, meanCoverage=means, groupMeans=groupMeans)


# ---------------------------------------------------------------------------------------------
# F statistics  (derfinderHelper::fstats.apply; normative restatement tests/testthat/test_Fstats.R:52-63)


def _residual_projector(M):
    M = np.asarray(M, dtype=np.float64)
    return np.eye(M.shape[0]) - M @ np.linalg.solve(M.T @ M, M.T)


def fstats_projector(Y, mod, mod0, adjustF=0.0):
    """F = ((rss0-rss1)/(df1-df0)) / (adjustF + rss1/(n-df1)) with rss = rowSums((Y %*% P)^2)
    (tests/testthat/test_Fstats.R:52-63; adjustF: R/calculateStats.R:24-26)."""
    Y = np.asarray(Y, dtype=np.float64)
    n = Y.shape[1]
    df1, df0 = mod.shape[1], mod0.shape[1]
    P1, P0 = _residual_projector(mod), _residual_projector(mod0)
    r0, r1 = Y @ P0, Y @ P1
    rss0 = (r0 * r0).sum(axis=1)
    rss1 = (r1 * r1).sum(axis=1)
    with np.errstate(divide="ignore", invalid="ignore"):
        return ((rss0 - rss1) / (df1 - df0)) / (adjustF + rss1 / (n - df1))


def calculate_stats(prep, mod, mod0, adjustF=0.0):
    """calculateStats(): per-chunk fstats.apply then concatenation (R/calculateStats.R:123-128).
    Chunking does not change per-base results; it is honoured anyway."""
    Y = prep["coverageProcessed"]
    if Y.shape[1] != mod.shape[0]:
        raise ValueError("The alternative model 'models$mod' is not compatible with the number of "
                         "samples in 'coveragePrep$coverageProcessed'.")
    out, off = [], 0
    for l in prep["mclapplyIndex"]:
        out.append(fstats_projector(Y[off:off + l], mod, mod0, adjustF))
        off += l
    return np.concatenate(out) if out else np.zeros(0)


def quantile_type7(x, prob):
    """R's default quantile() (type 7) with na.rm=TRUE."""
    x = np.sort(np.asarray(x, dtype=np.float64)[~np.isnan(x)])
    if len(x) == 0:  # quantile(c(NA, NA), na.rm = TRUE) is NA
        return float("nan")
    # stats:::quantile.default, type 7, in R's own operation order: index <- 1 + max(n - 1, 0) * probs;
    # lo <- floor(index); hi <- ceiling(index); h <- index - lo  (1-based ranks)
    index = 1.0 + (len(x) - 1) * prob
    lo = int(math.floor(index))
    hi = int(math.ceil(index))
    qs = x[lo - 1]
    if index > lo and x[hi - 1] != qs:
        g = index - lo
        qs = (1 - g) * qs + g * x[hi - 1]
    return qs


def calc_fstat_cutoff(cutoffType, cutoffFstat, fstats=None, mod=None, mod0=None):
    """.calcFstatCutoff (R/analyzeChr.R:312-328)."""
    if cutoffType == "empirical":
        return quantile_type7(fstats, cutoffFstat)
    if cutoffType == "theoretical":
        from scipy.stats import f as fdist
        n = mod.shape[0]
        return float(fdist.isf(cutoffFstat, mod.shape[1] - mod0.shape[1], n - mod.shape[1]))
    if cutoffType == "manual":
        return cutoffFstat
    raise ValueError(cutoffType)


# ---------------------------------------------------------------------------------------------
# .clusterMakerRle  (R/findRegions.R:445-475)


def cluster_maker_rle(position, maxGap=300, ranges=False):
    """position: (logical values, lengths).  ranges=False -> (ids, lengths) integer Rle over the
    TRUE positions; ranges=True -> (start, end) 1-based index-space IRanges tiling 1..sum(TRUE)."""
    v, l = rle_normalize(np.asarray(position[0]).astype(bool), position[1])
    ends = np.cumsum(l)
    starts = ends - l + 1
    st, en = starts[v], ends[v]  # TRUE runs (:448-451)
    if st.size == 0:
        z = np.zeros(0, dtype=np.int64)
        return (z, z)
    # reduce(ir, min.gapwidth = maxGap + 1): merge when gap <= maxGap (:454)
    gap = st[1:] - en[:-1] - 1
    new = np.concatenate(([True], gap > maxGap))
    grp = np.cumsum(new) - 1
    counts = np.zeros(int(grp[-1]) + 1, dtype=np.int64)
    np.add.at(counts, grp, en - st + 1)  # sum(Views(position, ir.red)) (:457-460)
    if ranges:  # :465-468
        csum = np.cumsum(counts)
        return (csum - counts + 1, csum)
    return (np.arange(1, counts.size + 1, dtype=np.int64), counts)


# ---------------------------------------------------------------------------------------------
# .getSegmentsRle  (R/findRegions.R:363-396)


def _true_runs(mask):
    """1-based inclusive (start, end) of maximal TRUE runs of a dense bool vector."""
    m = np.asarray(mask, dtype=bool)
    if m.size == 0:
        z = np.zeros(0, dtype=np.int64)
        return z, z
    d = np.diff(np.concatenate(([0], m.view(np.int8), [0])))
    st = np.flatnonzero(d == 1) + 1
    en = np.flatnonzero(d == -1)
    return st.astype(np.int64), en.astype(np.int64)


def sort_cutoff(cutoff):
    c = np.atleast_1d(np.asarray(cutoff, dtype=np.float64))
    if c.size > 2:
        raise ValueError("length(cutoff) <= 2 is not TRUE")
    if c.size == 1:
        c = np.array([-c[0], c[0]])
    c = np.sort(c)
    return float(c[0]), float(c[1])


def get_segments(x, cutoff):
    """slice(x, lower=U) / slice(x, upper=L): inclusive on both sides (IRanges defaults)."""
    L, U = sort_cutoff(cutoff)
    x = np.asarray(x, dtype=np.float64)
    with np.errstate(invalid="ignore"):
        return {"upIndex": _true_runs(x >= U), "dnIndex": _true_runs(x <= L)}


# ---------------------------------------------------------------------------------------------
# Views sums over a numeric Rle  (IRanges viewSums/viewMeans on RleViews)


def view_sums(x, starts, ends):
    """sum(Views(Rle(x), IRanges(starts, ends))): left-to-right double accumulation where a run of
    k identical adjacent values contributes `value * k` as ONE product (the Rle is iterated run by
    run).  Bit-level order is not pinned by any reference test; this is the natural
    reading of the IRanges C loop and reproduces the golden areas to 7e-12 (the residual being
    F-stat rounding).  The engine implements exactly this order."""
    x = np.asarray(x, dtype=np.float64)
    if len(starts) > 2000:  # same loop in C (oracle/fstat_oracle.c::oc_view_sums) for the larger cases
        from . import c_oracle
        return c_oracle.view_sums(x, starts, ends)
    out = np.zeros(len(starts), dtype=np.float64)
    vals, lens = rle_encode(x)
    run_end = np.cumsum(lens)  # 1-based inclusive end of each run
    run_start = run_end - lens + 1
    first = np.searchsorted(run_end, np.asarray(starts), side="left")
    for r, (s, e) in enumerate(zip(starts, ends)):
        k = int(first[r])
        acc = 0.0
        while k < len(vals) and run_start[k] <= e:
            lo = max(int(run_start[k]), int(s))
            hi = min(int(run_end[k]), int(e))
            acc = acc + float(vals[k]) * float(hi - lo + 1)
            k += 1
        out[r] = acc
    return out


# ---------------------------------------------------------------------------------------------
# findRegions  (R/findRegions.R:116-296)


def _split_by_segments(st, en, seg_start, seg_end):
    """disjoin(c(fcut, segmentIR)) pieces that overlap fcut (:205-214): every fcut run is cut at
    the segment boundaries it spans."""
    out_s, out_e = [], []
    k = np.searchsorted(seg_end, st, side="left")
    for a, b, j in zip(st, en, k):
        j = int(j)
        while a <= b:
            e = min(b, int(seg_end[j]))
            out_s.append(a)
            out_e.append(e)
            a = e + 1
            j += 1
    return np.asarray(out_s, dtype=np.int64), np.asarray(out_e, dtype=np.int64)


def find_regions(position, fstats, cutoff, maxClusterGap=300, maxRegionGap=0, segmentIR=None,
                 basic=False):
    """findRegions() without smoothing.  position: logical Rle (values, lengths) or None (allowed
    when segmentIR is given and basic=True).  Returns None when nothing passes (:162-165,185-193).

    basic=True  -> dict(area, width, stat) (up rows then dn rows, :273-278, 292)
    basic=False -> dict(start, end, value, area, indexStart, indexEnd, cluster, clusterL, name)
    """
    x = np.asarray(fstats, dtype=np.float64)
    if segmentIR is None:
        if position is None or not np.any(np.asarray(position[0]).astype(bool)):
            return None  # warning("Found no regions") :162-165
        segmentIR = cluster_maker_rle(position, maxGap=maxRegionGap, ranges=True)
    seg_start, seg_end = segmentIR
    segs = get_segments(x, cutoff)
    has = {k: len(v[0]) != 0 for k, v in segs.items()}
    if not any(has.values()):
        return None
    pos = None
    if not basic:
        pv, pl = np.asarray(position[0]).astype(bool), np.asarray(position[1], dtype=np.int64)
        pe = np.cumsum(pl)
        ps = pe - pl + 1
        # which(position) (:221) -- only evaluated at the needed indices
        kept_before = np.concatenate(([0], np.cumsum(np.where(pv, pl, 0))))

        def pos(idx):
            idx = np.asarray(idx, dtype=np.int64)
            tr = np.flatnonzero(pv)
            cum_true = kept_before[tr + 1]  # kept count through each TRUE run
            r = np.searchsorted(cum_true, idx, side="left")
            run = tr[r]
            return ps[run] + (idx - kept_before[run] - 1)

    parts = []
    for key, name in (("upIndex", "up"), ("dnIndex", "dn")):
        if not has[key]:
            continue
        st, en = _split_by_segments(*segs[key], seg_start, seg_end)
        sums = view_sums(x, st, en)
        width = en - st + 1
        with np.errstate(invalid="ignore", divide="ignore"):
            means = sums / width
        if basic:
            parts.append(dict(area=np.abs(sums), width=width, stat=means))
            continue
        gstart, gend = pos(st), pos(en)
        # clusters (:254-271)
        cov_s, cov_e = gstart, gend  # regions are disjoint and sorted: coverage() == their union
        gap = cov_s[1:] - cov_e[:-1] - 1
        # adjacent covered runs (gap == 0) are one TRUE run of regionPos; reduce merges gap<=maxClusterGap
        new = np.concatenate(([True], gap > maxClusterGap))
        region_cluster = np.cumsum(new)  # cluster id of the covered run each region forms
        gwidth = cov_e - cov_s + 1  # covered genomic bases contributed by each region
        # cluster Rle has one entry per covered genomic base; Views use cumulative *index* widths
        G = np.concatenate(([0], np.cumsum(gwidth)))
        Phi = np.concatenate(([0], np.cumsum(gwidth * region_cluster)))

        def prefix(t):  # sum of the first t entries of the cluster vector
            t = np.minimum(np.asarray(t, dtype=np.int64), G[-1])
            j = np.minimum(np.searchsorted(G, t, side="right") - 1, len(gwidth) - 1)
            return Phi[j] + (t - G[j]) * region_cluster[j]

        cw = np.cumsum(width)
        vsum = prefix(cw) - prefix(cw - width)
        with np.errstate(invalid="ignore"):
            clusterFinal = np.trunc(vsum.astype(np.float64) / width).astype(np.int64)  # as.integer(mean(clus))
        # clusterWidth <- tapply(pos.ir, clusterFinal, ...); clusterWidth[clusterFinal] indexes by
        # POSITION in the sorted unique ids (:266-271)
        uniq = np.unique(clusterFinal)
        cwid = np.array([gend[clusterFinal == u].max() - gstart[clusterFinal == u].min() + 1 for u in uniq],
                        dtype=np.float64)
        clusterL = np.full(len(clusterFinal), np.nan)
        ok = (clusterFinal >= 1) & (clusterFinal <= len(uniq))
        clusterL[ok] = cwid[clusterFinal[ok] - 1]
        parts.append(dict(start=gstart, end=gend, value=means, area=np.abs(sums), indexStart=st,
                          indexEnd=en, cluster=clusterFinal, clusterL=clusterL,
                          name=np.array([name] * len(st))))
    keys = parts[0].keys()
    return {k: np.concatenate([p[k] for p in parts]) for k in keys}


# ---------------------------------------------------------------------------------------------
# .calcPval  (R/calculatePvalues.R:432-441)   .calculateFWER  (R/mergeResults.R:380-386)


def calc_pval(areas, nullareas):
    nulls = np.sort(np.asarray(nullareas, dtype=np.float64))
    uniq, cnt = rle_encode(nulls)
    nullsum = len(nulls) - np.cumsum(cnt)
    table = np.concatenate(([len(nulls)], nullsum))
    idx = np.searchsorted(uniq, np.asarray(areas, dtype=np.float64), side="right")  # findInterval
    return (table[idx] + 1) / (len(nulls) + 1)


def calculate_fwer(cutoffFstatUsed, areaNull, permutation, nPermute, areaReg):
    areaNull = np.asarray(areaNull, dtype=np.float64)
    permutation = np.asarray(permutation)
    mx = np.empty(nPermute, dtype=np.float64)
    for b in range(1, nPermute + 1):
        sel = areaNull[permutation == b]
        mx[b - 1] = sel.max() if sel.size else cutoffFstatUsed
    areaReg = np.asarray(areaReg, dtype=np.float64)
    return ((mx[None, :] > areaReg[:, None]).sum(axis=1) + 1) / (nPermute + 1)


# ---------------------------------------------------------------------------------------------
# calculatePvalues  (R/calculatePvalues.R:157-430)


def calculate_pvalues(prep, mod, mod0, fstats, perm_idx, cutoff, maxRegionGap=0, maxClusterGap=300,
                      adjustF=0.0, significantCut=(0.05, 0.1), fstat_fn=None, fstat_idx_fn=None):
    """perm_idx: [B x n] 0-based permutation rows (== `sample(n) - 1` per seed).  qvalue() is a
    third-party spline fit that stays in R and is not restated.  `fstat_fn(Y, mod, mod0, adjustF)`
    lets tests swap the F-stat formulation (default: the projector definition)."""
    fstat_fn = fstat_fn or fstats_projector
    position, means, groupMeans = prep["position"], prep["meanCoverage"], prep["groupMeans"]
    Y = prep["coverageProcessed"]
    segmentIR = cluster_maker_rle(position, maxGap=maxRegionGap, ranges=True)  # :234-237
    regs = find_regions(position, fstats, cutoff, maxClusterGap=maxClusterGap,
                        maxRegionGap=maxRegionGap, segmentIR=segmentIR)  # :240-244
    if regs is None:
        return dict(regions=None, nullStats=None, nullWidths=None, nullPermutation=None)
    w = regs["indexEnd"] - regs["indexStart"] + 1
    regs["meanCoverage"] = view_sums(means, regs["indexStart"], regs["indexEnd"]) / w  # :254-255
    names = list(groupMeans.keys())
    gm = {}
    for g in names:  # :258-262
        gm[g] = view_sums(groupMeans[g], regs["indexStart"], regs["indexEnd"]) / w
        regs["mean" + str(g)] = gm[g]
    for g in names[1:]:  # :265-283
        with np.errstate(divide="ignore", invalid="ignore"):
            regs[f"log2FoldChange{g}vs{names[0]}"] = np.log2(gm[g] / gm[names[0]])
    nstat, nwidth, narea, nperm = [], [], [], []
    for i, idx in enumerate(np.asarray(perm_idx)):  # :306-358
        if fstat_idx_fn is not None:  # formulation that takes the row permutation itself
            f_b = fstat_idx_fn(Y, idx)
        else:
            f_b = fstat_fn(Y, mod[idx, :], mod0[idx, :], adjustF)
        rp = find_regions(None, f_b, cutoff, segmentIR=segmentIR, basic=True)
        if rp is not None:
            for _ in range(2):  # stored twice (:346-354)
                nstat.append(rp["stat"])
                nwidth.append(rp["width"])
                narea.append(rp["area"])
                nperm.append(np.full(len(rp["stat"]), i + 1, dtype=np.int64))
    cat = lambda xs, dt: (np.concatenate(xs) if xs else np.zeros(0, dtype=dt))
    nstat, nwidth = cat(nstat, np.float64), cat(nwidth, np.int64)
    narea, nperm = cat(narea, np.float64), cat(nperm, np.int64)
    order = np.argsort(-regs["area"], kind="stable")  # order(area, decreasing=TRUE) is stable (:368)
    regs = {k: v[order] for k, v in regs.items()}
    if len(nstat) > 0:
        regs["pvalues"] = calc_pval(regs["area"], narea)  # :380
        regs["significant"] = regs["pvalues"] < significantCut[0]
    else:
        regs["pvalues"] = np.full(len(order), np.nan)
        regs["significant"] = np.full(len(order), np.nan)
    return dict(regions=regs, nullStats=nstat, nullWidths=nwidth, nullPermutation=nperm,
                nullAreas=narea)


# ---------------------------------------------------------------------------------------------
# mergeResults: the genome-wide pass over the per-chromosome results  (R/mergeResults.R:199-328)


def merge_results(results, nPermute, cutoffFstatUsed=None, significantCut=(0.05, 0.10)):
    """`results`: ordered dict chr -> output of calculate_pvalues (the per-chromosome regions.Rdata).
    Returns dict(fullRegions, fullNullSummary) shaped like the objects mergeResults() saves; annotation, timing
    and q-values (qvalue package) are not restated."""
    chrs = [c for c in results if results[c]["regions"] is not None]
    regs = [results[c]["regions"] for c in chrs]
    keys = [k for k in regs[0] if all(k in r for r in regs)] if regs else []
    full = {k: np.concatenate([np.asarray(r[k]) for r in regs]) for k in keys}  # unlist(GRangesList(fullRegs)) :199
    full["seqnames"] = np.concatenate([np.repeat(c, len(r["area"])) for c, r in zip(chrs, regs)]) if regs else np.zeros(0, str)
    allc = list(results)
    stat = np.concatenate([np.asarray(results[c]["nullStats"], dtype=np.float64) if results[c]["nullStats"] is not None
                           else np.zeros(0) for c in allc]) if allc else np.zeros(0)
    width = np.concatenate([np.asarray(results[c]["nullWidths"], dtype=np.int64) if results[c]["nullWidths"] is not None
                            else np.zeros(0, np.int64) for c in allc]) if allc else np.zeros(0, np.int64)
    perm = np.concatenate([np.asarray(results[c]["nullPermutation"], dtype=np.int64)
                           if results[c]["nullPermutation"] is not None else np.zeros(0, np.int64) for c in allc]) \
        if allc else np.zeros(0, np.int64)
    how = [0 if results[c]["nullStats"] is None else len(results[c]["nullStats"]) for c in allc]
    if len(stat) > 0:  # :224-238
        area = np.abs(stat) * width  # :231 -- NOT the summed area of calculatePvalues
        o = np.argsort(-area, kind="stable")  # order(area, decreasing = TRUE)
        null = dict(stat=stat[o], width=width[o], chr=np.repeat(allc, how)[o], permutation=perm[o], area=area[o])
    else:
        null = None
    if cutoffFstatUsed is not None and null is not None and len(full.get("area", ())) > 0:  # :266-289
        full["fwer"] = calculate_fwer(cutoffFstatUsed, null["area"], null["permutation"], nPermute, full["area"])
        full["significantFWER"] = full["fwer"] < significantCut[0]
    if regs:
        o = np.argsort(-full["area"], kind="stable")  # :300
        full = {k: v[o] for k, v in full.items()}
        if null is not None:  # :302-327
            full["pvalues"] = calc_pval(full["area"], null["area"])
            full["significant"] = full["pvalues"] < significantCut[0]
    return dict(fullRegions=full if regs else None, fullNullSummary=null)


# ---------------------------------------------------------------------------------------------
# regionMatrix  (R/regionMatrix.R:170-283, R/getRegionCoverage.R:132-160)


def region_matrix(cols, cutoff=5, L=1, totalMapped=80e6, targetSize=80e6, maxRegionGap=0,
                  maxClusterGap=300, position=None):
    """One chromosome of regionMatrix(runFilter=TRUE).  cols: Rle columns of the raw coverage."""
    filt = filter_data(cols, cutoff=cutoff, index=position, filter="mean", totalMapped=totalMapped,
                       targetSize=targetSize, returnMean=True, returnCoverage=True)
    regs = find_regions(filt.position, filt.meanCoverage, cutoff, maxClusterGap=maxClusterGap,
                        maxRegionGap=maxRegionGap)
    if regs is None:
        return dict(regions=None, coverageMatrix=None)
    cov = filt.coverage
    R, n = len(regs["indexStart"]), cov.shape[1]
    mat = np.zeros((R, n), dtype=np.float64)
    for r in range(R):  # colSums of the region's rows (:260-262), rows added top to bottom
        blk = cov[regs["indexStart"][r] - 1:regs["indexEnd"][r]]
        acc = np.zeros(n, dtype=np.float64)
        for row in blk:
            acc = acc + row
        mat[r] = acc
    Lv = np.atleast_1d(np.asarray(L, dtype=np.float64))
    if Lv.size == 1:
        mat = mat / Lv[0]
    elif Lv.size == n:
        mat[:, n - 1] = mat[:, n - 1] / Lv[n - 1]  # `for (i in length(L))` only hits the LAST column (:267)
    return dict(regions=regs, coverageMatrix=mat, filtered=filt)


def rail_matrix(mean_cov, sample_cols, cutoff, L=1, maxClusterGap=300, totalMapped=None, targetSize=40e6,
                maxRegionGap=0):
    """One chromosome of railMatrix() (R/railMatrix.R:176-290) with the file contents in memory:
    mean_cov = (values, lengths) of the summary BigWig, sample_cols = per-sample Rle columns.  ERs of
    the mean above `cutoff` (:182-187), then per sample `regCov / mappedPerXM` (:286-288) and
    `sum(Views(regCov, ranges(regs))) / L` (:289) -- an Rle view sum, i.e. one product per run."""
    filt = filter_data([mean_cov], cutoff=cutoff, filter="mean", returnMean=True, returnCoverage=False)
    regs = None
    if len(filt.meanCoverage):
        regs = find_regions(filt.position, filt.meanCoverage, cutoff, maxClusterGap=maxClusterGap,
                            maxRegionGap=maxRegionGap)
    if regs is None:
        return dict(regions=None, coverageMatrix=None)
    n = len(sample_cols)
    Lv = np.broadcast_to(np.asarray(L, dtype=np.float64), (n,))
    mat = np.zeros((len(regs["start"]), n), dtype=np.float64)
    for s, (v, l) in enumerate(sample_cols):
        x = rle_decode(v, l).astype(np.float64)
        if totalMapped is not None and targetSize != 0:
            x = x / (np.broadcast_to(np.asarray(totalMapped, dtype=np.float64), (n,))[s] / targetSize)
        mat[:, s] = view_sums(x, regs["start"], regs["end"]) / Lv[s]
    return dict(regions=regs, coverageMatrix=mat)


def coverage_to_exon(dense, starts, ends, L=1):
    """R/coverageToExon.R:206-222 for one chromosome: colSums of every exon's rows of the dense
    [Lc x n] coverage, divided by L (scalar; a vector only divides the last column, the same
    `for (i in length(L))` quirk as regionMatrix)."""
    dense = np.asarray(dense)
    n = dense.shape[1]
    mat = np.zeros((len(starts), n), dtype=np.float64)
    for r, (a, b) in enumerate(zip(starts, ends)):
        acc = np.zeros(n, dtype=np.float64)
        for row in dense[a - 1:b]:
            acc = acc + row
        mat[r] = acc
    Lv = np.atleast_1d(np.asarray(L, dtype=np.float64))
    if Lv.size == 1:
        mat = mat / Lv[0]
    elif Lv.size == n:
        mat[:, n - 1] = mat[:, n - 1] / Lv[n - 1]
    return mat


# ---------------------------------------------------------------------------------------------
# Fixture helpers: sampleDepth / collapseFullCoverage / makeModels (host-side, tiny; only needed
# to rebuild the golden models)


def collapse_full_coverage(cols):
    """values/weights table per sample (R/collapseFullCoverage.R:73-99): unique run values in order
    of first appearance and the number of bases carrying them."""
    out = []
    for v, l in cols:
        v = np.asarray(v)
        l = np.asarray(l, dtype=np.int64)
        uniq = []
        seen = set()
        for x in v.tolist():
            if x not in seen:
                seen.add(x)
                uniq.append(x)
        w = [int(l[v == u].sum()) for u in uniq]
        out.append((np.asarray(uniq, dtype=np.float64), np.asarray(w, dtype=np.float64)))
    return out


def _wtd_quantile(x, w, p):
    """Hmisc::wtd.quantile(type='quantile', normwt=FALSE) for one probability."""
    o = np.argsort(x, kind="stable")
    x, w = x[o], w[o]
    cum = np.cumsum(w)
    n = cum[-1]
    order = 1 + (n - 1) * p
    low = max(math.floor(order), 1)
    high = min(low + 1, n)
    frac = order % 1
    # approx(cumsum(w), x, xout, method='constant', f=1, rule=2): right-continuous step lookup
    def look(t):
        k = int(np.searchsorted(cum, t, side="left"))
        return x[min(k, len(x) - 1)]
    a, b = look(low), look(high)
    return (1 - frac) * a + frac * b


def sample_depth(collapsed, probs=0.5, scalefac=32, nonzero=True):
    """sampleDepth() for a single probability (R/sampleDepth.R:98-121)."""
    out = []
    for v, w in collapsed:
        if nonzero:
            keep = v > 0
            v, w = v[keep], w[keep]
        if v.size == 0:
            out.append(math.log2(scalefac))
            continue
        q = _wtd_quantile(v, w, probs)
        sel = v <= q
        out.append(math.log2(float((v[sel] * w[sel]).sum()) + scalefac))
    return np.asarray(out)
