"""Host-side mirror of the reference's R interface for the hot path, on top of the C ABI.

The product boundary is the C ABI (include/derfinder_b200.h).  The reference's host language is R,
which this image does not have, so the operator-level interface is mirrored here in Python with the
reference's function names, argument meaning and error behaviour (the R glue in `r-package/` does
the same through `.Call`):

    filterData            R/filterData.R:89        preprocessCoverage   R/preprocessCoverage.R:97
    calculateStats        R/calculateStats.R:70    findRegions          R/findRegions.R:116
    calculatePvalues      R/calculatePvalues.R:157 regionMatrix         R/regionMatrix.R:124
    analyzeChr            R/analyzeChr.R:108       calcPval / calculateFWER / clusterMakerRle /
                                                   getSegmentsRle (internal helpers with tests)

Conventions: an Rle is a `(values, lengths)` pair of numpy arrays; a coverage "DataFrame" is a list
(or dict) of Rle columns, or a dense `[N x n]` array; regions come back as dicts of numpy arrays with
the reference's column names (1-based coordinates).  Everything numeric runs on the GPU; this file
only marshals.
"""
from __future__ import annotations

import ctypes as C
import time
import warnings
import weakref

import numpy as np

from . import _lib as L
from .rrng import permutation_indices

__all__ = ["Engine", "Coverage", "filterData", "preprocessCoverage", "calculateStats", "findRegions",
           "calculatePvalues", "regionMatrix", "analyzeChr", "calcPval", "calculateFWER", "clusterMakerRle",
           "getSegmentsRle", "calcFstatCutoff", "makeModels"]


# ---------------------------------------------------------------------------------------------
# engine / handles


class Engine:
    """One GPU context (stream, scratch pool, kernel timers)."""

    def __init__(self, device: int = 0, stream=None):
        lib = L.load()
        h = C.c_void_p()
        L.check(lib.dfb_create(int(device), C.c_void_p(stream) if stream else None, C.byref(h)))
        self.stream_handle = int(stream) if stream else None  # None: the library owns a private stream
        self.lib, self.h, self.device = lib, h, device
        self._covs = weakref.WeakSet()  # live Coverage handles: released before the context goes

    def close(self):
        for cov in list(getattr(self, "_covs", ())):
            cov.free()
        if getattr(self, "h", None):
            self.lib.dfb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        L.check(self.lib.dfb_sync(self.h))

    def set_option(self, key: str, value: float):
        L.check(self.lib.dfb_set_option(self.h, key.encode(), float(value)))

    def enable_timing(self, on=True, families=None):
        """Bracket kernel launches with CUDA events: every family, or only those listed (L.K_*)."""
        mask = 0
        for k in families or ():
            mask |= 1 << int(k)
        L.check(self.lib.dfb_enable_timing(self.h, mask if families else (1 if on else 0)))

    def reset_counters(self):
        self.lib.dfb_reset_counters(self.h)

    def launch_count(self) -> int:
        return int(self.lib.dfb_launch_count(self.h))

    def kernel_time(self, kernel: int):
        ms, n = C.c_double(), C.c_int64()
        L.check(self.lib.dfb_kernel_time(self.h, kernel, C.byref(ms), C.byref(n)))
        return ms.value, n.value


_default_engines: dict = {}


def default_engine(device: int = 0) -> Engine:
    if device not in _default_engines:
        _default_engines[device] = Engine(device)
    return _default_engines[device]


def _rle_norm(values, lengths):
    values = np.asarray(values)
    lengths = np.asarray(lengths, dtype=np.int64)
    keep = lengths > 0
    return values[keep], lengths[keep]


def _position_runs(position):
    """logical Rle -> (int32 run lengths with alternating values, value of the first run)."""
    v, l = _rle_norm(np.asarray(position[0]).astype(bool), position[1])
    if len(v) == 0:
        return np.zeros(0, dtype=np.int32), 0
    # merge equal neighbours so that values alternate
    new = np.concatenate(([True], v[1:] != v[:-1]))
    grp = np.cumsum(new) - 1
    out = np.zeros(int(grp[-1]) + 1, dtype=np.int64)
    np.add.at(out, grp, l)
    return out.astype(np.int32), int(v[0])


def _as_columns(data):
    """list/dict of Rle columns or dense [N x n] array -> ('rle', names, cols) | ('dense', names, array)."""
    if isinstance(data, Coverage):
        return "handle", data.names, data
    if isinstance(data, np.ndarray):
        if data.ndim != 2:
            raise ValueError("dense coverage must be a [N x n] array")
        return "dense", [f"V{i + 1}" for i in range(data.shape[1])], data
    if isinstance(data, dict):
        return "rle", list(data.keys()), list(data.values())
    return "rle", [f"V{i + 1}" for i in range(len(data))], list(data)


def _marshal_rle(cols):
    n = len(cols)
    vals, lens = [], []
    is_int = all(np.issubdtype(np.asarray(c[0]).dtype, np.integer) or np.asarray(c[0]).dtype == bool for c in cols)
    for v, l in cols:
        v, l = _rle_norm(v, l)
        vals.append(np.ascontiguousarray(v, dtype=np.int32 if is_int else np.float64))
        lens.append(np.ascontiguousarray(l, dtype=np.int32))
    vp = (C.c_void_p * n)(*[v.ctypes.data for v in vals])
    lp = (C.c_void_p * n)(*[l.ctypes.data for l in lens])
    nr = np.asarray([len(v) for v in vals], dtype=np.int64)
    total = int(lens[0].astype(np.int64).sum()) if n else 0
    return vals, lens, vp, lp, nr, (L.DFB_I32 if is_int else L.DFB_F64), total


class Coverage:
    """Device-resident filtered coverage of one chromosome (`dfb_cov`): the coverage DataFrame, its
    position index and, after preprocessCoverage, the log2 matrix, means and cached F-statistics."""

    def __init__(self, engine: Engine, handle, names=None):
        self.engine, self.h = engine, handle
        N, n, clen, dt, npr, G, hp = (C.c_int64(), C.c_int(), C.c_int64(), C.c_int(), C.c_int64(), C.c_int(),
                                      C.c_int())
        L.check(engine.lib.dfb_cov_info(handle, C.byref(N), C.byref(n), C.byref(clen), C.byref(dt), C.byref(npr),
                                        C.byref(G), C.byref(hp)))
        self.N, self.n, self.chr_len, self.dtype = N.value, n.value, clen.value, dt.value
        self._npr = npr.value
        self.names = list(names) if names is not None else [f"V{i + 1}" for i in range(self.n)]
        # mean vectors handed out by preprocessCoverage(): when calculatePvalues() gets these very objects
        # back it uses the device copies; any other vector is uploaded and used as given
        self._issued = {}
        engine._covs.add(self)

    def free(self):
        # a handle must not outlive its context: once the engine is closed the device memory has
        # gone with the stream's pool and the handle is simply dropped
        if self.h and getattr(self.engine, "h", None):
            self.engine.lib.dfb_cov_free(self.h)
        self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def __len__(self):
        return self.n

    @property
    def shape(self):
        return (self.N, self.n)

    @property
    def position(self):
        lens = np.zeros(self._npr, dtype=np.int32)
        first = C.c_int()
        L.check(self.engine.lib.dfb_cov_get_position(self.h, L.ptr(lens, C.c_int32), C.byref(first)))
        vals = (np.arange(self._npr) % 2 == 0) == bool(first.value)
        return vals, lens.astype(np.int64)

    def column(self, s: int):
        out = np.empty(self.N, dtype=np.int32 if self.dtype == L.DFB_I32 else np.float64)
        L.check(self.engine.lib.dfb_cov_get_column(self.h, int(s), out.ctypes.data_as(C.c_void_p)))
        return out

    def to_numpy(self):
        """[N x n] filtered coverage."""
        return np.stack([self.column(s) for s in range(self.n)], axis=1) if self.n else np.zeros((self.N, 0))

    def processed_column(self, s: int):
        out = np.empty(self.N, dtype=np.float64)
        L.check(self.engine.lib.dfb_cov_get_processed_column(self.h, int(s), L.ptr(out, C.c_double)))
        return out

    def processed(self):
        """[N x n] log2(x + scalefac) -- `coverageProcessed`."""
        return np.stack([self.processed_column(s) for s in range(self.n)], axis=1)

    def mean(self):
        out = np.empty(self.N, dtype=np.float64)
        L.check(self.engine.lib.dfb_cov_get_mean(self.h, L.ptr(out, C.c_double)))
        return out

    def group_mean(self, g: int):
        out = np.empty(self.N, dtype=np.float64)
        L.check(self.engine.lib.dfb_cov_get_group_mean(self.h, int(g), L.ptr(out, C.c_double)))
        return out


def _upload_filtered(engine, data, position, names=None):
    """Coverage handle from already-filtered data (dense [N x n] or Rle columns of length N)."""
    if position is None:
        raise ValueError("'coverageInfo$position' is NULL when it shouldn't be. Use a cutoff when loading the data "
                         "with loadCoverage() or fullCoverage(). Alternatively, you can specify 'colsubset' such that "
                         "preproccessCoverage() will run filterData() again.")
    kind, nm, cols = _as_columns(data)
    if kind == "handle":
        return cols
    pl, pf = _position_runs(position)
    h = C.c_void_p()
    if kind == "dense":
        N, n = cols.shape
        is_int = np.issubdtype(cols.dtype, np.integer)
        mat = np.ascontiguousarray(cols.T, dtype=np.int32 if is_int else np.float64)  # sample-major
        L.check(engine.lib.dfb_cov_from_dense(engine.h, n, N, L.DFB_I32 if is_int else L.DFB_F64,
                                              mat.ctypes.data_as(C.c_void_p), N, L.ptr(pl, C.c_int32), len(pl), pf,
                                              C.byref(h)))
    else:
        vals, lens, vp, lp, nr, dt, total = _marshal_rle(cols)
        L.check(engine.lib.dfb_cov_from_rle(engine.h, len(cols), dt, vp, lp, L.ptr(nr, C.c_int64), total,
                                            L.ptr(pl, C.c_int32), len(pl), pf, C.byref(h)))
    return Coverage(engine, h, names or nm)


# ---------------------------------------------------------------------------------------------
# filterData  (R/filterData.R:89-241)


def filterData(data, cutoff=None, index=None, filter="one", totalMapped=None, targetSize=80e6, returnMean=False,
               returnCoverage=True, colnames=None, engine=None, verbose=False, **kwargs):
    if filter not in ("one", "mean"):
        raise ValueError("filter %in% c(\"one\", \"mean\") is not TRUE")
    if kwargs.get("smoothMean"):
        raise NotImplementedError("smoothMean = TRUE is not supported by the B200 engine (bumphunter smoothing)")
    engine = engine or default_engine()
    kind, names, cols = _as_columns(data)
    if kind == "handle":
        raise TypeError("filterData() takes host Rle columns or a dense matrix; the data is already on the device")
    if kind == "dense":  # dense input: encode the columns (host marshalling of the caller's format)
        cols = [_dense_to_rle(cols[:, s]) for s in range(cols.shape[1])]
    n = len(cols)
    vals, lens, vp, lp, nr, dt, total = _marshal_rle(cols)
    mapped = None
    if totalMapped is not None and targetSize != 0:  # :119-130
        mapped = np.broadcast_to(np.asarray(totalMapped, dtype=np.float64) / float(targetSize), (n,)).copy()
    fk = L.FILTER_NONE if cutoff is None else (L.FILTER_ONE if filter == "one" else L.FILTER_MEAN)
    pl, pf = (None, 0)
    if index is not None:
        pl, pf = _position_runs(index)
    h = C.c_void_p()
    L.check(engine.lib.dfb_filter_rle(engine.h, n, dt, vp, lp, L.ptr(nr, C.c_int64), total, L.ptr(mapped, C.c_double),
                                      fk, float(cutoff if cutoff is not None else 0.0), L.ptr(pl, C.c_int32),
                                      0 if pl is None else len(pl), pf, C.byref(h)))
    cov = Coverage(engine, h, colnames or names)
    res = {}
    if returnCoverage:
        res["coverage"] = cov
    # cutoff = NULL: position = index (may be NULL) (:133-135)
    res["position"] = (index if cutoff is None else cov.position)
    if returnMean:
        res["meanCoverage"] = cov.mean()
    if not returnCoverage:
        cov.free()
    return res


def _dense_to_rle(x):
    x = np.asarray(x)
    if x.size == 0:
        return x, np.zeros(0, dtype=np.int64)
    brk = np.flatnonzero(x[1:] != x[:-1]) + 1
    st = np.concatenate(([0], brk))
    return x[st], np.diff(np.concatenate((st, [x.size])))


# ---------------------------------------------------------------------------------------------
# preprocessCoverage  (R/preprocessCoverage.R:97-260)


def preprocessCoverage(coverageInfo, groupInfo=None, cutoff=5, colsubset=None, lowMemDir=None, scalefac=32,
                       chunksize=5e6, mc_cores=1, engine=None, verbose=False, **kwargs):
    if not ("coverage" in coverageInfo and "position" in coverageInfo):
        raise ValueError("coverageInfo must have 'coverage' and 'position' (output of loadCoverage/filterData)")
    engine = engine or default_engine()
    coverage, position = coverageInfo["coverage"], coverageInfo["position"]
    means = coverageInfo.get("meanCoverage")
    kind, names, cols = _as_columns(coverage)
    ncols = cols.n if kind == "handle" else (cols.shape[1] if kind == "dense" else len(cols))
    if groupInfo is not None:
        expect = len(colsubset) if colsubset is not None else ncols
        if len(groupInfo) != expect:
            raise ValueError("length(groupInfo) must match the number of samples used")
    if colsubset is not None:  # re-filter on the subset (:145-160); colsubset is 1-based like R
        sub = [int(c) - 1 for c in colsubset]
        if kind == "handle":
            dense = cols.to_numpy()[:, sub]
        elif kind == "dense":
            dense = cols[:, sub]
        else:
            dense = None
        subdata = dense if dense is not None else [cols[c] for c in sub]
        filt = filterData(subdata, cutoff=cutoff, index=position, returnMean=True, engine=engine,
                          colnames=[names[c] for c in sub], **kwargs)
        cov, position = filt["coverage"], filt["position"]
        means = filt["meanCoverage"]  # the re-filtered mean replaces a pre-existing one (:157-159)
        cov._issued["mean"] = means
    else:
        cov = _upload_filtered(engine, coverage, position, names)
    if position is None:
        raise ValueError("'coverageInfo$position' is NULL when it shouldn't be.")
    levels, codes = None, None
    if groupInfo is not None:
        g = np.asarray(groupInfo)
        levels = sorted(set(g.tolist()))  # factor levels
        codes = np.asarray([levels.index(x) for x in g.tolist()], dtype=np.int32)
    L.check(engine.lib.dfb_preprocess(engine.h, cov.h, float(scalefac), L.ptr(codes, C.c_int32),
                                      0 if levels is None else len(levels)))
    nchunk = engine.lib.dfb_chunk_lengths(cov.N, float(chunksize) if chunksize else 0.0, int(mc_cores), None, 0)
    lens = np.zeros(nchunk, dtype=np.int64)
    engine.lib.dfb_chunk_lengths(cov.N, float(chunksize) if chunksize else 0.0, int(mc_cores), L.ptr(lens, C.c_int64),
                                 nchunk)
    groupMeans = {lev: cov.group_mean(i) for i, lev in enumerate(levels)} if levels else {}
    for i, lev in enumerate(levels or ()):
        cov._issued[("group", i)] = groupMeans[lev]
    if means is None:  # Reduce("+", coverage) / length(coverage) (:191-193): computed on the device
        means = cov.mean()
        cov._issued["mean"] = means
    else:
        means = np.asarray(means, dtype=np.float64)
        if len(means) != cov.N:
            raise ValueError("length(coverageInfo$meanCoverage) does not match the number of filtered bases")
    return dict(coverageProcessed=cov, mclapplyIndex=[int(x) for x in lens], position=position,
                meanCoverage=means, groupMeans=groupMeans)


# ---------------------------------------------------------------------------------------------
# calculateStats  (R/calculateStats.R:70-132)


def _models(models):
    if not ("mod" in models and "mod0" in models):
        raise ValueError("length(intersect(names(models), c(\"mod\", \"mod0\"))) == 2 is not TRUE")
    mod = np.asfortranarray(np.asarray(models["mod"], dtype=np.float64))
    mod0 = np.asfortranarray(np.asarray(models["mod0"], dtype=np.float64))
    return mod, mod0


def _fptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def calculateStats(coveragePrep, models, lowMemDir=None, adjustF=0.0, engine=None, **kwargs):
    for k in ("coverageProcessed", "mclapplyIndex", "position"):
        if k not in coveragePrep:
            raise ValueError("coveragePrep must be the output of preprocessCoverage()")
    mod, mod0 = _models(models)
    cov = coveragePrep["coverageProcessed"]
    if cov is None:
        raise ValueError("preprocessCoverage() was used with a non-null 'lowMemDir', so please specify 'lowMemDir'.")
    engine = engine or cov.engine
    if cov.n != mod.shape[0]:
        raise ValueError("The alternative model 'models$mod' is not compatible with the number of samples in "
                         "'coveragePrep$coverageProcessed'. Check the dimensions of the alternative model.")
    out = np.empty(cov.N, dtype=np.float64)
    L.check(engine.lib.dfb_fstats(engine.h, cov.h, _fptr(mod), mod.shape[1], _fptr(mod0), mod0.shape[1],
                                  float(adjustF), L.ptr(out, C.c_double)))
    return out


# ---------------------------------------------------------------------------------------------
# findRegions  (R/findRegions.R:116-296)


def _cutoff_pair(cutoff):
    c = np.atleast_1d(np.asarray(cutoff, dtype=np.float64))
    if c.size > 2:
        raise ValueError("length(cutoff) <= 2 is not TRUE")
    if c.size == 1:
        c = np.array([-c[0], c[0]])
    c = np.sort(c)
    return float(c[0]), float(c[1])


def _regions_dict(engine, rh, basic):
    lib = engine.lib
    R = int(lib.dfb_regions_count(rh))
    Rup = int(lib.dfb_regions_count_up(rh))
    value, area = np.empty(R), np.empty(R)
    ist, ien = np.empty(R, dtype=np.int32), np.empty(R, dtype=np.int32)
    if basic:
        L.check(lib.dfb_regions_get(rh, None, None, L.ptr(value, C.c_double), L.ptr(area, C.c_double),
                                    L.ptr(ist, C.c_int32), L.ptr(ien, C.c_int32), None, None))
        return dict(area=area, width=(ien - ist + 1).astype(np.int64), stat=value)
    st, en, cl = np.empty(R, dtype=np.int32), np.empty(R, dtype=np.int32), np.empty(R, dtype=np.int32)
    cll = np.empty(R)
    L.check(lib.dfb_regions_get(rh, L.ptr(st, C.c_int32), L.ptr(en, C.c_int32), L.ptr(value, C.c_double),
                                L.ptr(area, C.c_double), L.ptr(ist, C.c_int32), L.ptr(ien, C.c_int32),
                                L.ptr(cl, C.c_int32), L.ptr(cll, C.c_double)))
    order = np.empty(R, dtype=np.int32)
    L.check(lib.dfb_regions_get_order(rh, L.ptr(order, C.c_int32)))
    name = np.where(order < Rup, "up", "dn")
    return dict(start=st.astype(np.int64), end=en.astype(np.int64), value=value, area=area,
                indexStart=ist.astype(np.int64), indexEnd=ien.astype(np.int64), cluster=cl.astype(np.int64),
                clusterL=cll, name=name)


def findRegions(position=None, fstats=None, chr=None, oneTable=True, maxClusterGap=300, cutoff=None, segmentIR=None,
                smooth=False, weights=None, smoothFunction=None, basic=False, maxRegionGap=0, engine=None,
                verbose=False, **kwargs):
    if maxClusterGap < maxRegionGap:
        warnings.warn("'maxClusterGap' is less than 'maxRegionGap' which nullifies it's intended use.")
    if smooth:
        if basic:
            warnings.warn("Ignoring 'smooth' = TRUE since 'basic' = TRUE")
        else:
            raise NotImplementedError("smooth = TRUE is not supported by the B200 engine (bumphunter smoothers)")
    if not basic and segmentIR is None and position is None:
        raise ValueError("!is.null(position) is not TRUE")  # R/findRegions.R:141-144
    engine = engine or default_engine()
    x = L.f64(fstats)
    if segmentIR is None or position is not None:
        if position is None:
            raise ValueError("!is.null(position) is not TRUE")
        if segmentIR is not None:
            # segmentIR is the cache of .clusterMakerRle(position, maxRegionGap, ranges = TRUE)
            # (R/calculatePvalues.R:234-237): the engine derives the same segments from `position`
            st, en = (np.asarray(a, dtype=np.int64) for a in segmentIR)
            kept = int(np.asarray(position[1])[np.asarray(position[0]).astype(bool)].sum())
            if int((en - st + 1).sum()) != kept:
                raise ValueError("sum(width(segmentIR)) == sum(position) is not TRUE")
        pl, pf = _position_runs(position)
        if not np.any(np.asarray(position[0]).astype(bool)):
            warnings.warn("Found no regions")
            return None
    else:
        # basic = TRUE with the segments only (index space, R/findRegions.R:155-170): a position index whose TRUE
        # runs ARE the segments, one filtered-out base between neighbours
        if not basic:
            raise ValueError("!is.null(position) is not TRUE")
        st, en = (np.asarray(a, dtype=np.int64) for a in segmentIR)
        w = en - st + 1
        if len(w) == 0 or int(w.sum()) != len(x) or st[0] != 1 or np.any(st[1:] != en[:-1] + 1):
            raise ValueError("segmentIR must tile 1..length(fstats)")
        runs = np.empty(2 * len(w) - 1, dtype=np.int64)
        runs[0::2] = w
        runs[1::2] = 1
        vals = np.arange(len(runs)) % 2 == 0
        pl, pf = _position_runs((vals, runs))
        maxRegionGap = 0
    if cutoff is None:  # quantile(fstats, 0.99, na.rm = TRUE) (:118)
        q = C.c_double()
        L.check(engine.lib.dfb_quantile(engine.h, None, L.ptr(x, C.c_double), len(x), 0.99, C.byref(q)))
        cutoff = q.value
    lo, hi = _cutoff_pair(cutoff)
    rh = C.c_void_p()
    rc = L.check(engine.lib.dfb_find_regions_pos(engine.h, L.ptr(x, C.c_double), len(x), L.ptr(pl, C.c_int32),
                                                 len(pl), pf, lo, hi, int(maxRegionGap), int(maxClusterGap),
                                                 1 if basic else 0, C.byref(rh)))
    if rc == L.DFB_ENOREGIONS:
        return None
    try:
        return _regions_dict(engine, rh, basic)
    finally:
        engine.lib.dfb_regions_free(rh)


def getSegmentsRle(x, cutoff, engine=None):
    """.getSegmentsRle (R/findRegions.R:363-396): up/dn index ranges (1-based, inclusive)."""
    x = L.f64(x)
    pos = (np.array([True]), np.array([len(x)]))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        r = findRegions(pos, x, cutoff=cutoff, maxRegionGap=0, maxClusterGap=0, engine=engine)
    empty = (np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64))
    if r is None:
        return dict(upIndex=empty, dnIndex=empty)
    up = r["name"] == "up"
    return dict(upIndex=(r["indexStart"][up], r["indexEnd"][up]), dnIndex=(r["indexStart"][~up], r["indexEnd"][~up]))


def clusterMakerRle(position, maxGap=300, ranges=False):
    """.clusterMakerRle (R/findRegions.R:445-475)."""
    lib = L.load()
    pl, pf = _position_runs(position)
    k = lib.dfb_cluster_maker(L.ptr(pl, C.c_int32), len(pl), pf, int(maxGap), None, None, None, 0)
    cnt, st, en = (np.zeros(k, dtype=np.int64) for _ in range(3))
    lib.dfb_cluster_maker(L.ptr(pl, C.c_int32), len(pl), pf, int(maxGap), L.ptr(cnt, C.c_int64), L.ptr(st, C.c_int64),
                          L.ptr(en, C.c_int64), k)
    if ranges:
        return st, en
    return np.arange(1, k + 1, dtype=np.int64), cnt


# ---------------------------------------------------------------------------------------------
# p-value helpers


def calcPval(areas, nullareas, engine=None):
    """.calcPval (R/calculatePvalues.R:432-441)."""
    engine = engine or default_engine()
    a, nu = L.f64(areas), L.f64(nullareas)
    out = np.empty(len(a))
    L.check(engine.lib.dfb_calc_pval(engine.h, L.ptr(a, C.c_double), len(a), L.ptr(nu, C.c_double), len(nu), 1,
                                     L.ptr(out, C.c_double)))
    return out


def calculateFWER(cutoffFstatUsed, areaNull, permutation, nPermute, areaReg, engine=None):
    """.calculateFWER (R/mergeResults.R:380-386)."""
    engine = engine or default_engine()
    a, nu, pm = L.f64(areaReg), L.f64(areaNull), L.i32(permutation)
    out = np.empty(len(a))
    L.check(engine.lib.dfb_calc_fwer(engine.h, float(cutoffFstatUsed), L.ptr(nu, C.c_double), L.ptr(pm, C.c_int32),
                                     len(nu), int(nPermute), L.ptr(a, C.c_double), len(a), L.ptr(out, C.c_double)))
    return out


def calcFstatCutoff(cutoffType="theoretical", cutoffFstat=1e-8, fstats=None, models=None, engine=None):
    """.calcFstatCutoff (R/analyzeChr.R:312-328)."""
    if cutoffType == "empirical":
        engine = engine or default_engine()
        x = L.f64(fstats)
        q = C.c_double()
        L.check(engine.lib.dfb_quantile(engine.h, None, L.ptr(x, C.c_double), len(x), float(cutoffFstat), C.byref(q)))
        return q.value
    if cutoffType == "theoretical":
        from scipy.stats import f as fdist  # qf() stays on the host (one scalar), as in the R glue
        mod, mod0 = _models(models)
        n = mod.shape[0]
        return float(fdist.isf(cutoffFstat, mod.shape[1] - mod0.shape[1], n - mod.shape[1]))
    if cutoffType == "manual":
        return cutoffFstat
    raise ValueError("cutoffType must be 'empirical', 'theoretical' or 'manual'")


def makeModels(sampleDepths, testvars, adjustvars=None):
    """Minimal stand-in for makeModels() (R/makeModels.R:91-107) for two-group `testvars`:
    mod = [1, testvars == level2, sampleDepths, adjustvars...], mod0 = mod without the test column.
    The real makeModels stays in R (host-side n x p matrices)."""
    t = np.asarray(testvars)
    lev = sorted(set(t.tolist()))
    colsx = [np.ones(len(t))] + [(t == l) * 1.0 for l in lev[1:]] + [np.asarray(sampleDepths, dtype=np.float64)]
    if adjustvars is not None:
        a = np.asarray(adjustvars, dtype=np.float64)
        a = a.reshape(len(t), -1)
        colsx += [a[:, j] for j in range(a.shape[1])]
    mod = np.column_stack(colsx)
    keep = [0] + list(range(len(lev), mod.shape[1]))
    return dict(mod=mod, mod0=mod[:, keep])


# ---------------------------------------------------------------------------------------------
# calculatePvalues  (R/calculatePvalues.R:157-430)


def calculatePvalues(coveragePrep, models, fstats, nPermute=1, seeds=None, chr=None, cutoff=None,
                     significantCut=(0.05, 0.1), lowMemDir=None, smooth=False, weights=None, smoothFunction=None,
                     adjustF=0.0, maxRegionGap=0, maxClusterGap=300, sampleKind="Rejection", permIndex=None,
                     engine=None, verbose=False, **kwargs):
    if smooth:
        raise NotImplementedError("smooth = TRUE is not supported by the B200 engine (bumphunter smoothers)")
    for k in ("coverageProcessed", "mclapplyIndex", "position", "meanCoverage", "groupMeans"):
        if k not in coveragePrep:
            raise ValueError("coveragePrep must be the output of preprocessCoverage()")
    mod, mod0 = _models(models)
    if len(significantCut) != 2 or not all(0 <= s <= 1 for s in significantCut):
        raise ValueError("length(significantCut) == 2 & all(significantCut >= 0 & significantCut <= 1) is not TRUE")
    cov = coveragePrep["coverageProcessed"]
    if cov is None:
        raise ValueError("preprocessCoverage() was used with a non-null 'lowMemDir', so please specify 'lowMemDir'.")
    engine = engine or cov.engine
    n = mod.shape[0]
    if permIndex is None:
        if seeds is None:
            raise ValueError("seeds = NULL continues the R session's RNG stream; pass explicit seeds or permIndex")
        if nPermute != len(seeds):
            raise ValueError("nPermute == length(seeds) is not TRUE")
        permIndex = permutation_indices(list(seeds), n, sampleKind)  # set.seed(seeds[i]); sample(n) (:315-318)
    perm = L.i32(permIndex).reshape(-1, n)
    B = perm.shape[0]
    x = L.f64(fstats)
    if cutoff is None:
        cutoff = calcFstatCutoff("empirical", 0.99, fstats=x, engine=engine)
    lo, hi = _cutoff_pair(cutoff)
    rh, nh = C.c_void_p(), C.c_void_p()
    rc = L.check(engine.lib.dfb_calculate_pvalues(engine.h, cov.h, L.ptr(x, C.c_double), _fptr(mod), mod.shape[1],
                                                  _fptr(mod0), mod0.shape[1], float(adjustF), L.ptr(perm, C.c_int32),
                                                  B, lo, hi, int(maxRegionGap), int(maxClusterGap), C.byref(rh),
                                                  C.byref(nh), None))
    if rc == L.DFB_ENOREGIONS:  # :245-251
        return dict(regions=None, nullStats=None, nullWidths=None, nullPermutation=None)
    lib = engine.lib
    try:
        regs = _regions_dict(engine, rh, False)
        R = len(regs["start"])

        def view_means(vec, key, which):
            # mean(Views(vec, indexIR)) (:253-262) of the vector the caller passed: the device copy when
            # it is the object preprocessCoverage() returned for this handle, else the vector as given
            out = np.empty(R)
            if vec is cov._issued.get(key):
                L.check(lib.dfb_regions_view_means(engine.h, rh, cov.h, which, L.ptr(out, C.c_double)))
            else:
                v = L.f64(vec)
                L.check(lib.dfb_regions_view_means_host(engine.h, rh, L.ptr(v, C.c_double), len(v),
                                                        L.ptr(out, C.c_double)))
            return out

        regs["meanCoverage"] = view_means(coveragePrep["meanCoverage"], "mean", -1)  # :254-255
        names = list(coveragePrep["groupMeans"].keys())
        gm = {}
        for gi, g in enumerate(names):  # :258-262
            v = view_means(coveragePrep["groupMeans"][g], ("group", gi), gi)
            gm[g] = v
            regs["mean" + str(g)] = v
        for g in names[1:]:  # :265-283 (R glue computes this ratio/log2 on the R vectors)
            with np.errstate(divide="ignore", invalid="ignore"):
                regs[f"log2FoldChange{g}vs{names[0]}"] = np.log2(gm[g] / gm[names[0]])
        M = int(lib.dfb_nulls_count(nh))
        stat, area = np.empty(2 * M), np.empty(2 * M)
        width, pid = np.empty(2 * M, dtype=np.int32), np.empty(2 * M, dtype=np.int32)
        L.check(lib.dfb_nulls_get(nh, 1, L.ptr(stat, C.c_double), L.ptr(width, C.c_int32), L.ptr(area, C.c_double),
                                  L.ptr(pid, C.c_int32)))
        pv = np.empty(R)
        L.check(lib.dfb_regions_get_pvalues(rh, L.ptr(pv, C.c_double)))
        if M > 0:
            regs["pvalues"] = pv
            regs["significant"] = pv < significantCut[0]
        else:  # :408-419
            regs["pvalues"] = np.full(R, np.nan)
            regs["significant"] = np.full(R, np.nan)
        # qvalue::qvalue() is a third-party spline fit that stays in R
        regs["qvalues"] = np.full(R, np.nan)
        regs["significantQval"] = np.full(R, np.nan)
        return dict(regions=regs, nullStats=stat, nullWidths=width.astype(np.int64),
                    nullPermutation=pid.astype(np.int64), nullAreas=area)
    finally:
        lib.dfb_regions_free(rh)
        lib.dfb_nulls_free(nh)


# ---------------------------------------------------------------------------------------------
# mergeResults  (R/mergeResults.R:109-357) -- the genome-wide pass, without the file plumbing


def mergeResults(chrs=None, results=None, optionsStats=None, significantCut=(0.05, 0.10), cutoffFstatUsed=None,
                 fullFstats=None, prefix=None, engine=None, verbose=False, **kwargs):
    """Genome-wide pooling of per-chromosome results.  `results`: dict chr -> output of calculatePvalues() /
    analyzeChr()["regions"] (what the reference loads from `<prefix>/<chr>/regions.Rdata`, :170-175).  Returns the
    objects the reference saves -- `fullRegions` (regions of all chromosomes with `fwer`, `significantFWER`, pooled
    `pvalues`, `significant`, in decreasing-area order, :199, :266-327), `fullNullSummary` (stat, width, chr,
    permutation, area = abs(stat) * width in decreasing order, :216-238) and `optionsMerge` -- and, with `prefix`,
    writes them as .npz next to where the reference writes the .Rdata files.  The ranking (.calcPval) and the
    per-permutation maxima (.calculateFWER) run on the device; for chromosomes whose device handles are still alive
    `pool.merge_results` does the same over all GPUs in one library call.  Annotation (`genomicState`), timing and
    qvalue() stay in R."""
    if len(significantCut) != 2 or not all(0 <= s <= 1 for s in significantCut):
        raise ValueError("length(significantCut) == 2 & all(significantCut >= 0 & significantCut <= 1) is not TRUE")
    if results is None:
        raise ValueError("'results' (chr -> calculatePvalues() output) is required: the engine has no .Rdata reader")
    chrs = list(results) if chrs is None else [c for c in chrs]
    engine = engine or default_engine()
    have = [c for c in chrs if results[c].get("regions") is not None]
    regs = [results[c]["regions"] for c in have]
    full = None
    if regs:
        keys = [k for k in regs[0] if all(k in r for r in regs)]
        full = {k: np.concatenate([np.asarray(r[k]) for r in regs]) for k in keys}
        full["seqnames"] = np.concatenate([np.repeat(c, len(r["area"])) for c, r in zip(have, regs)])

    def cat(key, dt):
        parts = [np.asarray(results[c][key], dtype=dt) for c in chrs if results[c].get(key) is not None]
        return np.concatenate(parts) if parts else np.zeros(0, dtype=dt)

    stat, width, perm = cat("nullStats", np.float64), cat("nullWidths", np.int64), cat("nullPermutation", np.int64)
    how = [0 if results[c].get("nullStats") is None else len(results[c]["nullStats"]) for c in chrs]
    null = None
    if len(stat) > 0:
        area = np.abs(stat) * width  # :231
        o = np.argsort(-area, kind="stable")
        null = dict(stat=stat[o], width=width[o], chr=np.repeat(np.asarray(chrs, dtype=object), how)[o],
                    permutation=perm[o], area=area[o])
    if cutoffFstatUsed is None and optionsStats is not None:  # :241-263
        for k in ("cutoffType", "cutoffFstat", "models"):
            if k not in optionsStats:
                raise ValueError('all(c("cutoffType", "cutoffFstat", "models") %in% names(optionsStats)) is not TRUE')
        fs = None if fullFstats is None else np.concatenate([np.asarray(fullFstats[c], dtype=np.float64) for c in chrs])
        cutoffFstatUsed = calcFstatCutoff(optionsStats["cutoffType"], optionsStats["cutoffFstat"], fstats=fs,
                                          models=optionsStats["models"], engine=engine)
    if cutoffFstatUsed is not None and full is not None:
        if null is not None:  # :266-283
            if optionsStats is None or "nPermute" not in optionsStats:
                raise ValueError('"nPermute" %in% names(optionsStats) is not TRUE')
            full["fwer"] = calculateFWER(cutoffFstatUsed, null["area"], null["permutation"], optionsStats["nPermute"],
                                         full["area"], engine=engine)
            full["significantFWER"] = full["fwer"] < significantCut[0]
        elif verbose:
            warnings.warn("No null regions were found, so the FWER calculation step will be skipped.")
    if full is not None:
        o = np.argsort(-full["area"], kind="stable")  # :300
        full = {k: v[o] for k, v in full.items()}
        if null is not None:  # :302-327
            full["pvalues"] = calcPval(full["area"], null["area"], engine=engine)
            full["significant"] = full["pvalues"] < significantCut[0]
            full["qvalues"] = np.full(len(o), np.nan)  # qvalue::qvalue() stays in R
            full["significantQval"] = np.full(len(o), np.nan)
    optionsMerge = dict(chrs=chrs, significantCut=tuple(significantCut), cutoffFstatUsed=cutoffFstatUsed,
                        optionsStats=optionsStats)
    out = dict(fullRegions=full, fullNullSummary=null, optionsMerge=optionsMerge)
    if prefix is not None:
        import os
        os.makedirs(prefix, exist_ok=True)
        if full is not None:
            np.savez(os.path.join(prefix, "fullRegions.npz"), **{k: np.asarray(v) for k, v in full.items()})
        if null is not None:
            np.savez(os.path.join(prefix, "fullNullSummary.npz"), **{k: np.asarray(v).astype(str) if k == "chr" else v
                                                                      for k, v in null.items()})
    return out


# ---------------------------------------------------------------------------------------------
# regionMatrix  (R/regionMatrix.R:124-283)


def regionMatrix(fullCov, cutoff=5, L_=None, totalMapped=80e6, targetSize=80e6, runFilter=True, returnBP=True,
                 maxRegionGap=0, maxClusterGap=300, engine=None, verbose=False, **kwargs):
    """fullCov: {chr: Rle columns | dense array | {'coverage':..., 'position':...}}.  `L_` is the
    reference's `L` (read length); the trailing underscore only avoids the module alias."""
    L_read = kwargs.pop("L", L_)
    if cutoff is None:
        raise ValueError("!is.null(cutoff) is not TRUE")
    if not isinstance(fullCov, dict) or not fullCov:
        raise ValueError("!is.null(names(fullCov)) is not TRUE")
    if L_read is None:
        raise ValueError("argument \"L\" is missing, with no default")
    if not runFilter and np.any(np.asarray(totalMapped) != targetSize):
        raise ValueError("When using 'runFilter' = FALSE the arguments 'totalMapped' and 'targetSize' are not used. "
                         "If you have not normalized the data, please do so before running "
                         "regionMatrix(runFilter = FALSE).")
    engine = engine or default_engine()
    out = {}
    for chrom, covInfo in fullCov.items():
        if runFilter:
            if isinstance(covInfo, dict) and "coverage" in covInfo and "position" in covInfo:
                data, index = covInfo["coverage"], covInfo["position"]
            else:
                data, index = covInfo, None
            filt = filterData(data, cutoff=cutoff, index=index, filter="mean", totalMapped=totalMapped,
                              targetSize=targetSize, returnMean=True, returnCoverage=True, engine=engine)
        else:
            if not all(k in covInfo for k in ("coverage", "position", "meanCoverage")):
                raise ValueError("When 'runFilter' = FALSE, all the elements of 'fullCov' are expected to be the "
                                 "output from filterData(..., returnMean=TRUE) with some non-NULL 'cutoff'.")
            filt = covInfo
        cov = filt["coverage"]
        if cov.N == 0 or not np.any(filt["position"][0]):
            out[chrom] = dict(regions=None, coverageMatrix=None, **({"bpCoverage": None} if returnBP else {}))
            continue
        lo, hi = _cutoff_pair(cutoff)
        mean = L.f64(filt["meanCoverage"])
        rh = C.c_void_p()
        pl, pf = _position_runs(filt["position"])
        rc = L.check(engine.lib.dfb_find_regions_pos(engine.h, L.ptr(mean, C.c_double), len(mean),
                                                     L.ptr(pl, C.c_int32), len(pl), pf, lo, hi, int(maxRegionGap),
                                                     int(maxClusterGap), 0, C.byref(rh)))
        if rc == L.DFB_ENOREGIONS:
            out[chrom] = dict(regions=None, coverageMatrix=None, **({"bpCoverage": None} if returnBP else {}))
            continue
        try:
            regs = _regions_dict(engine, rh, False)
            R = len(regs["start"])
            Lv = np.atleast_1d(np.asarray(L_read, dtype=np.float64))
            if Lv.size not in (1, cov.n):
                warnings.warn("Invalid 'L' value so it won't be used. It has to either be a integer/numeric vector "
                              "of length 1 or length equal to the number of samples.")
                Lv = np.zeros(0)
            mat = np.empty((cov.n, R), dtype=np.float64)  # column-major [R x n]
            L.check(engine.lib.dfb_region_matrix_r(engine.h, cov.h, rh, L.ptr(Lv, C.c_double) if Lv.size else None,
                                                   int(Lv.size), L.ptr(mat, C.c_double)))
            res = dict(regions=regs, coverageMatrix=mat.T.copy())
            if returnBP:  # per-region [width x n] blocks: plain row gathers of the filtered coverage
                dense = cov.to_numpy()
                res["bpCoverage"] = [dense[a - 1:b] for a, b in zip(regs["indexStart"], regs["indexEnd"])]
            out[chrom] = res
        finally:
            engine.lib.dfb_regions_free(rh)
    return out


# ---------------------------------------------------------------------------------------------
# coverageToExon  (R/coverageToExon.R:112-241): per exon x sample coverage sums / L -- the same
# segmented column-sum primitive as regionMatrix (K6), driven by annotation ranges


def coverageToExon(fullCov, exons, L_=None, returnType="raw", engine=None, verbose=False, **kwargs):
    """fullCov: {chr: Rle columns | dense [Lc x n]} of UNFILTERED coverage (fullCoverage() output);
    exons: {chr: (start, end)} 1-based inclusive, disjoint within a chromosome (the reduced exons of a
    genomic state, which the reference processes strand by strand, R/coverageToExon.R:131-150).
    Returns {chr: [nExons x n] matrix} in the order given; `L_` is the reference's `L`."""
    L_read = kwargs.pop("L", L_)
    if L_read is None:
        raise ValueError("argument \"L\" is missing, with no default")
    if returnType not in ("raw", "rpkm"):
        raise ValueError("length(returnType) == 1 & returnType %in% c('raw', 'rpkm') is not TRUE")
    engine = engine or default_engine()
    out, nsamp, msum = {}, None, None
    for chrom, data in fullCov.items():
        kind, names, cols = _as_columns(data)
        if kind == "dense":
            chr_len, n = cols.shape
            runsum = cols.astype(np.float64)
            runsum = np.where(np.vstack([np.ones((1, n), bool), cols[1:] != cols[:-1]]), runsum, 0.0).sum(axis=0)
        else:
            n = len(cols)
            chr_len = int(np.asarray(cols[0][1], dtype=np.int64).sum())
            runsum = np.array([np.asarray(_rle_norm(v, l)[0], dtype=np.float64).sum() for v, l in cols])
        nsamp = n if nsamp is None else nsamp
        msum = runsum if msum is None else msum + runsum  # colSums(Mchr): sum of RUN VALUES (:155-160)
        if chrom not in exons:
            continue
        st = L.i32(exons[chrom][0])
        en = L.i32(exons[chrom][1])
        if np.any(st < 1) or np.any(en > chr_len) or np.any(en < st):
            raise ValueError(f"exon ranges of {chrom} fall outside the chromosome")
        position = (np.array([True]), np.array([chr_len]))
        cov = _upload_filtered(engine, data, position, names)
        Lv = np.atleast_1d(np.asarray(L_read, dtype=np.float64))
        if Lv.size not in (1, n):
            warnings.warn("Invalid 'L' value so it won't be used. It has to either be a integer/numeric vector of "
                          "length 1 or length equal to the number of samples.")
            Lv = np.zeros(0)
        R = len(st)
        mat = np.empty((n, R), dtype=np.float64)  # column-major [R x n]
        L.check(engine.lib.dfb_region_matrix(engine.h, cov.h, L.ptr(st, C.c_int32), L.ptr(en, C.c_int32), R,
                                             L.ptr(Lv, C.c_double) if Lv.size else None, int(Lv.size),
                                             L.ptr(mat, C.c_double)))
        out[chrom] = mat.T.copy()
        cov.free()
    if returnType == "rpkm":  # :154-161
        Lv = np.asarray(L_read, dtype=np.float64)
        M = msum / Lv / 1e6
        for chrom, mat in out.items():
            width = (np.asarray(exons[chrom][1], dtype=np.float64) - np.asarray(exons[chrom][0], dtype=np.float64) + 1)
            out[chrom] = mat / (width[:, None] / 1000.0) / M[None, :]
    return out


# ---------------------------------------------------------------------------------------------
# railMatrix  (R/railMatrix.R:115-290) with the BigWig contents handed over in memory


def _rle_window(values, lengths, starts, ends):
    """The Rle restricted to the union of sorted, disjoint, 1-based inclusive ranges -- what
    `import(sampleFile, selection = reduce(regs))` reads (R/railMatrix.R:273-276); O(runs) host
    marshalling so that only the bases inside regions ever travel to the device."""
    values, lengths = _rle_norm(values, lengths)
    run_end = np.cumsum(lengths)
    run_start = run_end - lengths + 1
    first = np.searchsorted(run_end, starts, side="left")
    last = np.searchsorted(run_start, ends, side="right") - 1
    cnt = last - first + 1
    tot = int(cnt.sum())
    reg = np.repeat(np.arange(len(starts)), cnt)
    idx = np.arange(tot) - np.repeat(np.cumsum(cnt) - cnt, cnt) + np.repeat(first, cnt)
    lo = np.maximum(run_start[idx], np.asarray(starts, dtype=np.int64)[reg])
    hi = np.minimum(run_end[idx], np.asarray(ends, dtype=np.int64)[reg])
    return values[idx], hi - lo + 1


def railMatrix(chrs, summaryCov, sampleCov, L_=None, cutoff=None, maxClusterGap=300, totalMapped=None,
               targetSize=40e6, maxRegionGap=0, engine=None, verbose=False, deviceBytes=2 ** 31, **kwargs):
    """railMatrix() on coverage already in memory (reading the BigWig files stays with the caller):
    summaryCov {chr: (values, lengths)} is the content of the mean BigWig `summaryFiles[chr]`,
    sampleCov {chr: Rle columns | dense [Lc x n]} the per-sample BigWig contents.  Per chromosome
    (.railMatrixChr :176-234): ERs of the summary above `cutoff` (mean filter + findRegions), then per
    sample `sum(Views(cov / mappedPerXM, regions)) / L` (.railMatChrRegionCov :264-290).  Regions are
    processed in groups bounded by `deviceBytes` of expanded coverage (the reference's `chunksize`
    groups of 1000 regions; results do not depend on it, tests/testthat/test_ER_level.R:156)."""
    L_read = kwargs.pop("L", L_)
    chrs = [chrs] if isinstance(chrs, str) else list(chrs)
    if len(chrs) != len(summaryCov):
        raise ValueError("length(chrs) == length(summaryFiles) is not TRUE")
    if cutoff is None:
        raise ValueError("!is.null(cutoff) is not TRUE")
    if L_read is None:
        raise ValueError("argument \"L\" is missing, with no default")
    if kwargs.get("smoothMean"):
        raise NotImplementedError("smoothMean = TRUE is not supported by the B200 engine (bumphunter smoothing)")
    engine = engine or default_engine()
    out = {}
    for chrom in chrs:
        summary = summaryCov[chrom]
        if isinstance(summary, str):  # the mean BigWig itself (loadCoverage(files = summaryFile), :181)
            from .bigwig import BigWigFile
            bw = BigWigFile(summary)
            summary = bw.read_rle(chrom)
            bw.close()
        filt = filterData([summary], cutoff=cutoff, filter="mean", returnMean=True, returnCoverage=False,
                          engine=engine)
        regs = None
        if len(filt["meanCoverage"]):
            regs = findRegions(position=filt["position"], fstats=filt["meanCoverage"], chr=chrom,
                               maxClusterGap=maxClusterGap, cutoff=cutoff, maxRegionGap=maxRegionGap, engine=engine)
        if regs is None:  # :190-192
            out[chrom] = dict(regions=None, coverageMatrix=None)
            continue
        samples = sampleCov[chrom]
        if isinstance(samples, (list, tuple)) and len(samples) and all(isinstance(f, str) for f in samples):
            # sample BigWig files: the data sections overlapping the regions are inflated on the host and decoded
            # on the device (dfb_cov_from_bigwig), :271-289
            out[chrom] = dict(regions=regs, coverageMatrix=_rail_from_bigwig(engine, chrom, list(samples), regs, L_read,
                                                                            totalMapped, targetSize))
            continue
        kind, names, cols = _as_columns(samples)
        if kind == "dense":
            cols = [_dense_to_rle(cols[:, s]) for s in range(cols.shape[1])]
        n = len(cols)
        Lv = np.atleast_1d(np.asarray(L_read, dtype=np.float64))
        if Lv.size not in (1, n):
            raise ValueError("'L' has to be of length 1 or one value per sample")
        st = np.asarray(regs["start"], dtype=np.int64)
        en = np.asarray(regs["end"], dtype=np.int64)
        R = len(st)
        by_pos = np.argsort(st, kind="stable")  # windows are cut in genomic order
        width = (en - st + 1)[by_pos]
        mat = np.empty((R, n), dtype=np.float64)
        per_base = 8 * n
        r0 = 0
        while r0 < R:
            r1 = r0 + max(1, int(np.searchsorted(np.cumsum(width[r0:]) * per_base, deviceBytes, side="right")))
            sel = by_pos[r0:r1]
            win = [_rle_window(v, l, st[sel], en[sel]) for v, l in cols]
            cov = filterData(win, cutoff=None, totalMapped=totalMapped, targetSize=targetSize, engine=engine,
                             colnames=names)["coverage"]
            ce = np.cumsum(width[r0:r1])
            cs = (ce - width[r0:r1] + 1).astype(np.int32)
            ce = ce.astype(np.int32)
            part = np.empty((n, r1 - r0), dtype=np.float64)  # column-major [R x n]
            L.check(engine.lib.dfb_view_sums(engine.h, cov.h, L.ptr(cs, C.c_int32), L.ptr(ce, C.c_int32), r1 - r0,
                                             L.ptr(Lv, C.c_double), int(Lv.size), L.ptr(part, C.c_double)))
            mat[sel] = part.T
            cov.free()
            r0 = r1
        out[chrom] = dict(regions=regs, coverageMatrix=mat)
    return out


def _rail_from_bigwig(engine, chrom, files, regs, L_read, totalMapped, targetSize):
    """Region x sample matrix straight from the samples' BigWig files (.railMatChrRegionCov, R/railMatrix.R:264-290)."""
    from .bigwig import BigWigFile
    n = len(files)
    Lv = np.atleast_1d(np.asarray(L_read, dtype=np.float64))
    if Lv.size not in (1, n):
        raise ValueError("'L' has to be of length 1 or one value per sample")
    st = np.asarray(regs["start"], dtype=np.int64)
    en = np.asarray(regs["end"], dtype=np.int64)
    R = len(st)
    by_pos = np.argsort(st, kind="stable")
    rs, re = st[by_pos].astype(np.int32), en[by_pos].astype(np.int32)
    mapped = None
    if totalMapped is not None and targetSize != 0:  # :151-155
        mapped = np.broadcast_to(np.asarray(totalMapped, dtype=np.float64) / float(targetSize), (n,)).copy()
    blobs, offs, ids = [], [], np.zeros(n, dtype=np.uint32)
    for s, f in enumerate(files):
        bw = BigWigFile(f)
        blob, off, cid = bw.sections(chrom, rs, re)
        bw.close()
        blobs.append(np.frombuffer(blob, dtype=np.uint8) if len(blob) else np.zeros(1, dtype=np.uint8))
        offs.append(off)
        ids[s] = cid
    bp = (C.c_void_p * n)(*[b.ctypes.data for b in blobs])
    op = (C.c_void_p * n)(*[o.ctypes.data for o in offs])
    nsec = np.asarray([len(o) - 1 for o in offs], dtype=np.int64)
    h = C.c_void_p()
    L.check(engine.lib.dfb_cov_from_bigwig(engine.h, n, bp, op, L.ptr(nsec, C.c_int64), L.ptr(ids, C.c_uint32),
                                           L.ptr(rs, C.c_int32), L.ptr(re, C.c_int32), R, L.ptr(mapped, C.c_double),
                                           C.byref(h)))
    cov = Coverage(engine, h, [str(f) for f in files])
    try:
        width = (re.astype(np.int64) - rs + 1)
        ce = np.cumsum(width)
        cs = (ce - width + 1).astype(np.int32)
        ce = ce.astype(np.int32)
        part = np.empty((n, R), dtype=np.float64)  # column-major [R x n]
        L.check(engine.lib.dfb_view_sums(engine.h, cov.h, L.ptr(cs, C.c_int32), L.ptr(ce, C.c_int32), R,
                                         L.ptr(Lv, C.c_double), int(Lv.size), L.ptr(part, C.c_double)))
    finally:
        cov.free()
    mat = np.empty((R, n), dtype=np.float64)
    mat[by_pos] = part.T
    return mat


# ---------------------------------------------------------------------------------------------
# analyzeChr  (R/analyzeChr.R:108-310) without annotation / file output


def analyzeChr(chr, coverageInfo, models, cutoffPre=5, cutoffFstat=1e-8, cutoffType="theoretical", nPermute=1,
               seeds=None, groupInfo=None, txdb=None, writeOutput=False, runAnnotation=False, lowMemDir=None,
               smooth=False, engine=None, **kwargs):
    if runAnnotation or writeOutput:
        raise NotImplementedError("annotation (bumphunter/TxDb) and .Rdata output stay in the R glue")
    if cutoffType not in ("empirical", "theoretical", "manual"):
        raise ValueError("cutoffType %in% c('empirical', 'theoretical', 'manual') is not TRUE")
    if seeds is None:
        raise ValueError("pass explicit seeds (the date-based default depends on Sys.Date())")
    timeinfo = {"init": time.time()}
    prep = preprocessCoverage(coverageInfo, groupInfo=groupInfo, cutoff=cutoffPre, engine=engine, **kwargs)
    timeinfo["prepData"] = time.time()
    fstats = calculateStats(prep, models, engine=engine, **{k: v for k, v in kwargs.items() if k == "adjustF"})
    timeinfo["calculateStats"] = time.time()
    cut = calcFstatCutoff(cutoffType, cutoffFstat, fstats=fstats, models=models, engine=engine)
    regions = calculatePvalues(prep, models, fstats, nPermute=nPermute, seeds=seeds, chr=chr, cutoff=cut,
                               engine=engine, **kwargs)
    timeinfo["calculatePvalues"] = time.time()
    opts = dict(models=models, cutoffPre=cutoffPre, cutoffFstat=cutoffFstat, cutoffType=cutoffType, nPermute=nPermute,
                seeds=seeds, groupInfo=groupInfo, cutoffFstatUsed=cut)
    return dict(timeinfo=timeinfo, optionsStats=opts, coveragePrep=prep, fstats=fstats, regions=regions,
                annotation=None)
