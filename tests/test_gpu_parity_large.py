"""GPU parity at the parameters of BASELINE.json configs[2] / configs[3] (60 and 100 samples, p = 6,
p0 = 5, 1000 permutations, stringent theoretical cutoffs), on slices the oracle finishes in seconds,
plus the multi-rank pooling kernels (several ranks' records on one device) against .calcPval /
.calculateFWER of the oracle on the concatenated nulls (R/mergeResults.R:216-235, 300-307, 380-386).

Bars as in test_gpu_parity.py: regions, widths, permutation ids and p-value counts bit-exact; F
statistics bit-exact against the C oracle of the same arithmetic and <= 1e-6 relative against the
reference's projector definition.
"""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import c_oracle  # noqa: E402
from oracle import derfinder_oracle as O  # noqa: E402
from test_gpu_parity import (F_RTOL, assert_pipeline_equal, debug_screen, oracle_idx_fn, random_perms,  # noqa: E402
                             run_both, synth)


@pytest.fixture(scope="module")
def dfb():
    import derfinder_b200 as m
    return m


@pytest.fixture(scope="module")
def eng(dfb):
    e = dfb.Engine(0)
    yield e
    e.close()


# ---------------------------------------------------------------------------------------------
# C3 / C4 parameters on slices: n = 60 / 100, p = 6, p0 = 5, B = 1000


@pytest.mark.parametrize("n,N", [(60, 12000), (100, 6000)])
def test_c3_c4_parameters_vs_oracle(dfb, eng, n, N):
    from derfinder_b200.rrng import permutation_indices
    rng = np.random.default_rng(3000 + n)
    cov, position, group, models = synth(rng, N, n, p_extra=3, seg=[N // 3, 77, N - N // 3 - 77])
    assert models["mod"].shape[1] == 6 and models["mod0"].shape[1] == 5
    B = 1000
    perm = permutation_indices(range(1, B + 1), n, "Rejection")  # the relabellings R would draw for seeds 1:B
    for alpha, need_nulls in ((1e-6, False), (1e-3, True)):
        cut = dfb.calcFstatCutoff("theoretical", alpha, models=models)  # qf(alpha, 1, n - 6, lower.tail = FALSE)
        res, ores = run_both(dfb, eng, cov, position, group, models, perm, cut, cgap=3000)
        assert ores["regions"] is not None
        if need_nulls:
            assert len(ores["nullStats"]) > 100
        assert_pipeline_equal(res, ores)
        assert eng.kernel_time(3)[1] > 0  # the tcgen05 path ran


@pytest.mark.parametrize("n,extra,N,B", [(60, 3, 5000, 60), (100, 3, 3000, 40), (24, 2, 8000, 50)])
def test_permuted_nulls_vs_projector_definition(dfb, eng, n, extra, N, B):
    """Independent of the engine's own basis: null statistics of permuted models computed with the
    reference's residual-projector definition (tests/testthat/test_Fstats.R:52-63 on mod[idx, ],
    mod0[idx, ], R/calculatePvalues.R:321-329) must give the same null regions, and region means
    within the F-statistic tolerance."""
    rng = np.random.default_rng(4100 + n)
    cov, position, group, models = synth(rng, N, n, p_extra=extra)
    perm = random_perms(rng, B, n)
    cut = dfb.calcFstatCutoff("theoretical", 0.01, models=models)
    prep = dfb.preprocessCoverage(dict(coverage=cov, position=position), groupInfo=group, engine=eng)
    F = dfb.calculateStats(prep, models, engine=eng)
    res = dfb.calculatePvalues(prep, models, F, permIndex=perm, cutoff=cut, engine=eng)
    oprep = O.preprocess_coverage(cov, position, groupInfo=group)
    Y = oprep["coverageProcessed"]
    # per-base check first: a base whose projector F is within tolerance of the cutoff may legitimately fall
    # on either side; the region comparison below is only meaningful when no base is that close
    fn = oracle_idx_fn(eng, models["mod"], models["mod0"])
    close = 0
    for b in range(B):
        fp = c_oracle.fstats_projector(Y, models["mod"][perm[b]], models["mod0"][perm[b]])
        fe = fn(Y, perm[b])
        ok = np.isfinite(fp) & (fp > 1e-3)
        assert np.max(np.abs(fe[ok] - fp[ok]) / fp[ok]) < F_RTOL
        close += int(np.sum(np.abs(fp[ok] - cut) <= 2 * F_RTOL * cut))
    ores = O.calculate_pvalues(oprep, models["mod"], models["mod0"], F, perm, cut, fstat_fn=c_oracle.fstats_projector)
    assert len(ores["nullStats"]) > 0
    if close == 0:
        assert np.array_equal(res["nullWidths"], ores["nullWidths"])
        assert np.array_equal(res["nullPermutation"], ores["nullPermutation"])
        assert np.max(np.abs(res["nullStats"] - ores["nullStats"]) / ores["nullStats"]) < F_RTOL
        assert np.max(np.abs(res["nullAreas"] - ores["nullAreas"]) / ores["nullAreas"]) < F_RTOL


def _device_workload(dfb, name, seed):
    import torch
    import bench
    from derfinder_b200.rrng import permutation_indices
    cfg = bench.WORKLOADS[name]
    device = torch.device("cuda", 0)
    work = bench.make_workload(cfg, seed, torch=torch, device=device)
    perm = permutation_indices(range(1, cfg["B"] + 1), cfg["n"], "Rejection")
    return cfg, work, perm, torch, bench


@pytest.mark.parametrize("name", ["c3s", "c4s"])
def test_screen_exact_at_c3_c4_size(dfb, name):
    """dfb_debug_screen on the scaled C3 / C4 workloads (N = 2.5 M x 60 samples, N = 1 M x 100 samples,
    1000 permutations): the dense FP64 kernel and the tcgen05 screen + exact refinement (both the
    two-kernel and the fused form) must set identical mask bits -- the error bound of the screen is
    validated at the sample counts the headline configs use."""
    from derfinder_b200 import _lib as L
    cfg, work, perm, torch, bench = _device_workload(dfb, name, 20140923 + 3)
    e2 = dfb.Engine(0, stream=torch.cuda.current_stream().cuda_stream)
    try:
        pl = np.ascontiguousarray(work["pos_len"], dtype=np.int32)
        h = C.c_void_p()
        L.check(e2.lib.dfb_cov_view_dense_dev(e2.h, cfg["n"], cfg["N"], L.DFB_I32, C.c_void_p(work["cov"].data_ptr()),
                                              cfg["N"], L.ptr(pl, C.c_int32), len(pl), work["pos_first"], C.byref(h)))
        L.check(e2.lib.dfb_preprocess(e2.h, h, 32.0, None, 0))
        mod, mod0 = work["mod"], work["mod0"]
        B = 200  # the dense FP64 reference scan is the expensive side
        for alpha in (1e-6, 1e-3):
            cut = bench.theoretical_cutoff(alpha, mod, mod0)
            out = [C.c_int64() for _ in range(4)]
            L.check(e2.lib.dfb_debug_screen(e2.h, h, mod.ctypes.data_as(L.c_f64p), mod.shape[1],
                                            mod0.ctypes.data_as(L.c_f64p), mod0.shape[1], 0.0,
                                            L.ptr(np.ascontiguousarray(perm[:B]), C.c_int32), B, float(cut),
                                            *[C.byref(o) for o in out]))
            dense, screen, cand, diff = (o.value for o in out)
            assert diff == 0, (name, alpha, diff)
            assert dense == screen and cand >= dense
            if alpha == 1e-3:
                assert dense > 1000
        e2.lib.dfb_cov_free(h)
    finally:
        e2.close()


def test_c2_slice_vs_oracle(dfb, eng):
    """A 1e5-base slice of the configs[1] workload (12 samples, 100 permutations, qf(0.05, 1, 9)) end to
    end against the oracle."""
    import bench
    from derfinder_b200.rrng import permutation_indices
    cfg = dict(bench.WORKLOADS["c2"], chr_len=1_000_000, N=100_000)
    work = bench.make_workload(cfg, 20140923 + 2)
    cov = np.ascontiguousarray(work["cov"].T)
    pl = work["pos_len"].astype(np.int64)
    position = ((np.arange(len(pl)) % 2 == 0) == bool(work["pos_first"]), pl)
    models = dict(mod=np.array(work["mod"]), mod0=np.array(work["mod0"]))
    perm = permutation_indices(range(1, cfg["B"] + 1), cfg["n"], "Rejection")
    cut = bench.theoretical_cutoff(cfg["alpha"], work["mod"], work["mod0"])
    group = np.where(work["group"] == 0, "a", "b")
    res, ores = run_both(dfb, eng, cov, position, group, models, perm, cut, cgap=3000)
    assert len(ores["regions"]["start"]) > 100 and len(ores["nullStats"]) > 10000
    assert_pipeline_equal(res, ores)


# ---------------------------------------------------------------------------------------------
# multi-rank pooling kernels with world > 1, all "ranks" on one device


def _chromosome_result(dfb, eng, rng, N, n, models, perm, cut):
    """One rank's chromosome through the C ABI, handles kept: (cov, regions, nulls) or (cov, None, None)."""
    from derfinder_b200 import _lib as L
    cov, position, group, _ = synth(rng, N, n, p_extra=models["mod"].shape[1] - 3)
    mat = np.ascontiguousarray(cov.T, dtype=np.int32)
    pl = np.ascontiguousarray(position[1], dtype=np.int32)
    h = C.c_void_p()
    L.check(eng.lib.dfb_cov_from_dense(eng.h, n, N, L.DFB_I32, mat.ctypes.data_as(C.c_void_p), N, L.ptr(pl, C.c_int32),
                                       len(pl), int(position[0][0]), C.byref(h)))
    L.check(eng.lib.dfb_preprocess(eng.h, h, 32.0, None, 0))
    mod, mod0 = np.asfortranarray(models["mod"]), np.asfortranarray(models["mod0"])
    L.check(eng.lib.dfb_fstats(eng.h, h, mod.ctypes.data_as(L.c_f64p), mod.shape[1], mod0.ctypes.data_as(L.c_f64p),
                               mod0.shape[1], 0.0, None))
    rh, nh = C.c_void_p(), C.c_void_p()
    rc = L.check(eng.lib.dfb_calculate_pvalues(eng.h, h, None, mod.ctypes.data_as(L.c_f64p), mod.shape[1],
                                               mod0.ctypes.data_as(L.c_f64p), mod0.shape[1], 0.0,
                                               L.ptr(perm, C.c_int32), perm.shape[0], -cut, cut, 0, 300, C.byref(rh),
                                               C.byref(nh), None))
    if rc == L.DFB_ENOREGIONS:
        return h, None, None
    return h, rh, nh


def _fetch(eng, rh, nh):
    from derfinder_b200 import _lib as L
    lib = eng.lib
    R, M = int(lib.dfb_regions_count(rh)), int(lib.dfb_nulls_count(nh))
    area, order = np.empty(R), np.empty(R, dtype=np.int32)
    L.check(lib.dfb_regions_get(rh, None, None, None, L.ptr(area, C.c_double), None, None, None, None))
    L.check(lib.dfb_regions_get_order(rh, L.ptr(order, C.c_int32)))
    ns, na = np.empty(M), np.empty(M)
    nw, npm = np.empty(M, dtype=np.int32), np.empty(M, dtype=np.int32)
    if M:
        L.check(lib.dfb_nulls_get(nh, 0, L.ptr(ns, C.c_double), L.ptr(nw, C.c_int32), L.ptr(na, C.c_double),
                                  L.ptr(npm, C.c_int32)))
    return dict(R=R, M=M, area=area, order=order, nstat=ns, narea=na, nwidth=nw, nperm=npm)


@pytest.mark.parametrize("world,sizes,empty_rank", [(2, [9000, 4000], None), (3, [5000, 12000, 3000], 1),
                                                    (8, [3000, 7000, 2500, 4000, 6000, 2000, 5000, 3500], 5)])
@pytest.mark.parametrize("merge_area", [False, True])
def test_pool_kernels_world_gt_1(dfb, eng, world, sizes, empty_rank, merge_area):
    import torch
    from derfinder_b200 import _lib as L
    rng = np.random.default_rng(500 + world)
    n, B = 14, 24
    _, _, _, models = synth(rng, 2000, n, p_extra=1)
    perm = random_perms(rng, B, n)
    perm[3] = perm[2]  # a repeated relabelling: identical null areas on purpose (ties across the pool)
    cut = dfb.calcFstatCutoff("theoretical", 0.05, models=models)
    ranks = []
    for w in range(world):
        this_cut = 1e12 if w == empty_rank else cut  # nothing passes on that chromosome: regions = NULL
        ranks.append(_chromosome_result(dfb, eng, rng, sizes[w], n, models, perm, this_cut))
    try:
        info = [None if rh is None else _fetch(eng, rh, nh) for _, rh, nh in ranks]
        assert sum(i is not None for i in info) >= 2
        if empty_rank is not None:
            assert info[empty_rank] is None
        dev = torch.device("cuda", 0)
        lib = eng.lib
        ma = 1 if merge_area else 0
        key = (lambda i: np.abs(i["nstat"]) * i["nwidth"]) if merge_area else (lambda i: i["narea"])
        pool_area = np.concatenate([key(i) for i in info if i is not None])
        pool_perm = np.concatenate([i["nperm"] for i in info if i is not None])
        maxR = max(i["R"] for i in info if i is not None)
        for rcap in (maxR + 17, maxR):
            records = torch.empty((world, 2 + B + rcap), device=dev, dtype=torch.float64)
            for w, (_, rh, nh) in enumerate(ranks):
                L.check(lib.dfb_pool_local_dev(eng.h, rh, nh, B, ma, rcap, C.c_void_p(records[w].data_ptr())))
            eng.sync()
            head = records[:, :2].cpu().numpy()
            assert list(head[:, 0]) == [0 if i is None else i["R"] for i in info]
            assert list(head[:, 1]) == [0 if i is None else i["M"] for i in info]
            merged = torch.empty(world * rcap, device=dev, dtype=torch.float64)
            src = torch.empty(world * rcap, device=dev, dtype=torch.int32)
            total_hist = torch.zeros(world * rcap + 1, device=dev, dtype=torch.int64)
            for w, (_, rh, nh) in enumerate(ranks):  # every rank's histogram, then their sum (the all-reduce)
                hist = torch.empty(world * rcap + 1, device=dev, dtype=torch.int64)
                L.check(lib.dfb_pool_hist_dev(eng.h, nh, ma, C.c_void_p(records.data_ptr()), world, B, rcap,
                                              C.c_void_p(merged.data_ptr()), C.c_void_p(src.data_ptr()),
                                              C.c_void_p(hist.data_ptr())))
                eng.sync()
                total_hist += hist
            torch.cuda.synchronize()
            assert int(total_hist.sum().item()) == len(pool_area)
            for w, (_, rh, nh) in enumerate(ranks):
                if info[w] is None:
                    L.check(lib.dfb_pool_rank_dev(eng.h, None, C.c_void_p(records.data_ptr()), world, w, B, rcap,
                                                  C.c_void_p(merged.data_ptr()), C.c_void_p(src.data_ptr()),
                                                  C.c_void_p(total_hist.data_ptr()), cut, None, None))
                    continue
                R = info[w]["R"]
                p_dev = torch.full((R,), -1.0, device=dev, dtype=torch.float64)
                f_dev = torch.full((R,), -1.0, device=dev, dtype=torch.float64)
                L.check(lib.dfb_pool_rank_dev(eng.h, rh, C.c_void_p(records.data_ptr()), world, w, B, rcap,
                                              C.c_void_p(merged.data_ptr()), C.c_void_p(src.data_ptr()),
                                              C.c_void_p(total_hist.data_ptr()), cut, C.c_void_p(p_dev.data_ptr()),
                                              C.c_void_p(f_dev.data_ptr())))
                eng.sync()
                order = info[w]["order"]
                want_p = O.calc_pval(info[w]["area"], np.repeat(pool_area, 2))  # every null is stored twice
                want_f = O.calculate_fwer(cut, pool_area, pool_perm, B, info[w]["area"])
                assert np.array_equal(p_dev.cpu().numpy()[order], want_p), (world, w, rcap)
                assert np.array_equal(f_dev.cpu().numpy()[order], want_f), (world, w, rcap)
        # a capacity below some rank's region count: the record still reports the true R (the host logic
        # of pool.genomewide_pvalues repeats the exchange when it sees R > rcap)
        small = max(1, maxR // 2)
        rec = torch.empty(2 + B + small, device=dev, dtype=torch.float64)
        big = max(range(world), key=lambda w: -1 if info[w] is None else info[w]["R"])
        L.check(lib.dfb_pool_local_dev(eng.h, ranks[big][1], ranks[big][2], B, ma, small, C.c_void_p(rec.data_ptr())))
        eng.sync()
        assert int(rec[0].item()) == maxR > small
    finally:
        for h, rh, nh in ranks:
            if rh is not None:
                eng.lib.dfb_regions_free(rh)
                eng.lib.dfb_nulls_free(nh)
            eng.lib.dfb_cov_free(h)


def test_preprocess_colsubset_refreshes_mean(dfb, eng, golden):
    """preprocessCoverage(colsubset = ...) replaces a pre-existing meanCoverage with the re-filtered
    mean (R/preprocessCoverage.R:145-160); calculatePvalues uses the meanCoverage / groupMeans vectors it
    is handed (R/calculatePvalues.R:222-262)."""
    info = dict(coverage=golden.cov, position=golden.position,
                meanCoverage=golden.cov.astype(np.float64).mean(axis=1))  # stale: all 31 samples
    sub = [1, 2, 3, 4, 5, 6]
    grp = np.array(["a", "a", "a", "b", "b", "b"])
    ps = dfb.preprocessCoverage(info, groupInfo=grp, cutoff=1, colsubset=sub, chunksize=1e3, engine=eng)
    cols = [O.rle_encode(golden.cov[:, s - 1]) for s in sub]
    ds = O.filter_data(cols, 1, golden.position, returnMean=True)
    assert len(ps["meanCoverage"]) == ds.coverage.shape[0] < golden.cov.shape[0]
    assert np.array_equal(ps["meanCoverage"], ds.meanCoverage)
    # a caller-supplied mean vector is the one summarised per region
    rng = np.random.default_rng(12)
    cov, position, group, models = synth(rng, 6000, 10)
    prep = dfb.preprocessCoverage(dict(coverage=cov, position=position), groupInfo=group, engine=eng)
    F = dfb.calculateStats(prep, models, engine=eng)
    perm = random_perms(rng, 4, 10)
    base = dfb.calculatePvalues(prep, models, F, permIndex=perm, cutoff=2.0, engine=eng)
    mine = dict(prep)
    mine["meanCoverage"] = rng.uniform(0, 50, size=len(F))
    mine["groupMeans"] = {k: v * 2.0 for k, v in prep["groupMeans"].items()}
    got = dfb.calculatePvalues(mine, models, F, permIndex=perm, cutoff=2.0, engine=eng)
    r = got["regions"]
    w = r["indexEnd"] - r["indexStart"] + 1
    assert np.array_equal(r["meanCoverage"], O.view_sums(mine["meanCoverage"], r["indexStart"], r["indexEnd"]) / w)
    for g in prep["groupMeans"]:
        assert np.array_equal(r["mean" + g], O.view_sums(mine["groupMeans"][g], r["indexStart"], r["indexEnd"]) / w)
    assert np.array_equal(r["area"], base["regions"]["area"])
    assert not np.array_equal(r["meanCoverage"], base["regions"]["meanCoverage"])


# ---------------------------------------------------------------------------------------------
# dfb_merge_results: the one-call genome-wide pass (single rank here; tests/test_multi_gpu.py runs it over NCCL)


@pytest.mark.parametrize("merge_area", [False, True])
def test_merge_results_single_rank(dfb, eng, merge_area):
    from derfinder_b200 import pool
    rng = np.random.default_rng(909)
    n, B = 14, 24
    _, _, _, models = synth(rng, 2000, n, p_extra=1)
    perm = random_perms(rng, B, n)
    perm[5] = perm[4]
    cut = dfb.calcFstatCutoff("theoretical", 0.05, models=models)
    sizes = [9000, 2500, 6000, 4000]
    chroms = [_chromosome_result(dfb, eng, rng, sizes[c], n, models, perm, 1e12 if c == 1 else cut) for c in range(4)]
    try:
        info = [None if rh is None else _fetch(eng, rh, nh) for _, rh, nh in chroms]
        assert info[1] is None
        key = (lambda i: np.abs(i["nstat"]) * i["nwidth"]) if merge_area else (lambda i: i["narea"])
        pool_area = np.concatenate([key(i) for i in info if i is not None])
        pool_perm = np.concatenate([i["nperm"] for i in info if i is not None])
        outs, total = pool.merge_results(eng, [(rh, nh) for _, rh, nh in chroms], B, cut, merge_area=merge_area)
        assert total == len(pool_area)
        for c, i in enumerate(info):
            if i is None:
                assert len(outs[c][0]) == 0
                continue
            assert np.array_equal(outs[c][0], O.calc_pval(i["area"], np.repeat(pool_area, 2))), c
            assert np.array_equal(outs[c][1], O.calculate_fwer(cut, pool_area, pool_perm, B, i["area"])), c
        # nothing to pool at all
        outs, total = pool.merge_results(eng, [], B, cut)
        assert outs == [] and total == 0
    finally:
        for h, rh, nh in chroms:
            if rh is not None:
                eng.lib.dfb_regions_free(rh)
                eng.lib.dfb_nulls_free(nh)
            eng.lib.dfb_cov_free(h)


def test_quantile_radix_select(dfb, eng):
    """quantile(x, p, na.rm = TRUE) type 7 by radix select (no sort): negative values, signed zeros, infinities,
    ties, NaNs, one / two values, nothing but NaN (R/findRegions.R:118)."""
    rng = np.random.default_rng(77)
    cases = [rng.normal(size=100003), np.round(rng.normal(size=50000), 1), np.array([3.5]), np.array([2.0, -7.0]),
             np.concatenate((rng.normal(size=999), [np.inf, -np.inf, 0.0, -0.0, np.inf])),
             np.abs(rng.standard_cauchy(size=200001)) * 1e-300, np.full(7, np.nan), -np.abs(rng.gamma(2.0, size=4097))]
    x = rng.normal(size=30000)
    x[rng.random(30000) < 0.3] = np.nan
    cases.append(x)
    for x in cases:
        for p in (0.0, 1e-9, 0.25, 0.5, 0.99, 0.999999, 1.0, 0.123456):
            got = dfb.calcFstatCutoff("empirical", p, fstats=x, engine=eng)
            want = O.quantile_type7(x, p)
            assert (got == want) or (np.isnan(got) and np.isnan(want)), (len(x), p, got, want)


@pytest.mark.parametrize("kind,compress", [("bedGraph", True), ("bedGraph", False), ("varStep", True), ("fixedStep", True)])
def test_rail_matrix_from_bigwig(dfb, eng, tmp_path, kind, compress):
    """railMatrix() reading the mean and the sample BigWig files (R/railMatrix.R:181, 271-289): data sections
    are decoded on the device; result bit-identical to the oracle on the same coverage held in memory, and to
    the engine's own in-memory path, for the three BigWig section types, with and without library-size
    normalisation."""
    from bigwig_writer import write_bigwig
    rng = np.random.default_rng({"bedGraph": 1, "varStep": 2, "fixedStep": 3}[kind] + 10 * compress)
    n, Lc = 5, 30000

    def make_items():
        if kind == "bedGraph":
            items, pos = [], int(rng.integers(0, 50))
            while pos < Lc - 200:
                l, g = int(rng.integers(1, 80)), int(rng.integers(0, 40)) * int(rng.random() < 0.3)
                pos += g
                items.append((pos, min(pos + l, Lc), float(np.float32(rng.poisson(6.0) * 0.75))))
                pos += l
            return items, ("bedGraph", items)
        if kind == "varStep":
            span = 25
            starts = np.sort(rng.choice(np.arange(0, Lc - span, span), size=400, replace=False))
            pts = [(int(s), float(np.float32(rng.poisson(5.0) * 1.5))) for s in starts]
            return [(s, s + span, v) for s, v in pts], ("varStep", (span, pts))
        start, step, span = 100, 40, 30
        vals = [float(np.float32(rng.poisson(5.0) * 0.5)) for _ in range((Lc - 200) // step)]
        return [(start + i * step, start + i * step + span, v) for i, v in enumerate(vals)], ("fixedStep", (start, step, span, vals))

    dense = np.zeros((Lc, n))
    files = []
    for s in range(n):
        items, (k, payload) = make_items()
        for a, b, v in items:
            dense[a:b, s] = v
        path = str(tmp_path / f"s{s}.bw")
        write_bigwig(path, {"chr21": (Lc, k, payload), "chrZ": (500, "bedGraph", [(0, 10, 1.0)])}, items_per_section=53,
                     compress=compress)
        files.append(path)
    acc = np.zeros(Lc)
    for s in range(n):
        acc = acc + dense[:, s]
    mean = acc / n
    mean_items = []
    v, l = O.rle_encode(mean)
    pos = 0
    for vv, ll in zip(v, l):
        if vv != 0:
            mean_items.append((pos, pos + int(ll), float(vv)))
        pos += int(ll)
    # the mean track is stored in float32: use the values the file will give back
    mean_items = [(a, b, float(np.float32(x))) for a, b, x in mean_items]
    mpath = str(tmp_path / "mean.bw")
    write_bigwig(mpath, {"chr21": (Lc, "bedGraph", mean_items)}, items_per_section=101, compress=compress)
    mean32 = np.zeros(Lc)
    for a, b, x in mean_items:
        mean32[a:b] = x
    mean_rle = O.rle_encode(mean32)
    cols = [O.rle_encode(dense[:, s]) for s in range(n)]
    total_mapped = rng.integers(30_000_000, 90_000_000, size=n).astype(np.float64)
    for Lval, tm in ((76, None), (np.arange(1, n + 1) * 25.0, total_mapped)):
        got = dfb.railMatrix(["chr21"], {"chr21": mpath}, {"chr21": files}, L_=Lval, cutoff=2.2, maxClusterGap=3000,
                             totalMapped=tm, engine=eng)["chr21"]
        want = O.rail_matrix(mean_rle, cols, 2.2, L=Lval, maxClusterGap=3000, totalMapped=tm)
        assert len(want["regions"]["start"]) > 20
        for k in ("start", "end", "cluster", "area", "value"):
            assert np.array_equal(got["regions"][k], want["regions"][k]), k
        assert np.array_equal(got["coverageMatrix"], want["coverageMatrix"])
        mem = dfb.railMatrix(["chr21"], {"chr21": mean_rle}, {"chr21": cols}, L_=Lval, cutoff=2.2, maxClusterGap=3000,
                             totalMapped=tm, engine=eng)["chr21"]
        assert np.array_equal(got["coverageMatrix"], mem["coverageMatrix"])


# ---------------------------------------------------------------------------------------------
# log2 matrix not materialised (integer coverage read through the log2 table) == materialised


@pytest.mark.parametrize("n,extra,N,B,alpha", [(12, 0, 30000, 40, 0.05), (60, 3, 9000, 200, 1e-3)])
def test_lazy_log2_matrix_equals_materialised(dfb, n, extra, N, B, alpha):
    """dfb_preprocess leaves the log2 matrix un-materialised for integer coverage: the observed-F kernel and
    the refinement of the null scan read the int32 coverage through the host-built log2 table.  Same bits as
    the materialised path everywhere (F, regions, nulls, p-values), coverageProcessed still readable, and a
    coverage value beyond the context table falls back to the materialised matrix."""
    rng = np.random.default_rng(1234 + n)
    cov, position, group, models = synth(rng, N, n, p_extra=extra, seg=[N // 3, N - N // 3])
    cov[7, 0] = 5000  # beyond the shared-memory head of the table, inside the context table
    perm = random_perms(rng, B, n)
    cut = dfb.calcFstatCutoff("theoretical", alpha, models=models)
    out = {}
    for lazy in (1, 0):
        e = dfb.Engine(0)
        try:
            e.set_option("lazy_y", lazy)
            e.set_option("null2", 2)  # the un-materialised path belongs to the second-generation null scan
            prep = dfb.preprocessCoverage(dict(coverage=cov, position=position), groupInfo=group, engine=e)
            has = C.c_int(0)
            e.lib.dfb_cov_info(prep["coverageProcessed"].h, None, None, None, None, None, None, C.byref(has))
            assert has.value == 1
            F = dfb.calculateStats(prep, models, engine=e)
            res = dfb.calculatePvalues(prep, models, F, permIndex=perm, cutoff=cut, engine=e)
            Y = prep["coverageProcessed"].processed()
            out[lazy] = (np.asarray(F), res, Y, np.asarray(prep["meanCoverage"]))
        finally:
            e.close()
    assert np.array_equal(out[1][0], out[0][0])
    assert np.array_equal(out[1][2], out[0][2])
    assert np.array_equal(out[1][2], O.log2_libm(cov.astype(np.float64) + 32.0))
    assert np.array_equal(out[1][3], out[0][3])
    assert_pipeline_equal(out[1][1], out[0][1])
    e = dfb.Engine(0)
    try:
        oprep = O.preprocess_coverage(cov, position, groupInfo=group)
        ores = O.calculate_pvalues(oprep, models["mod"], models["mod0"], out[1][0], perm, cut,
                                   fstat_idx_fn=oracle_idx_fn(e, models["mod"], models["mod0"], 0.0))
        assert_pipeline_equal(out[1][1], ores)
    finally:
        e.close()


def test_lazy_log2_matrix_out_of_table(dfb, eng):
    rng = np.random.default_rng(99)
    cov, position, group, models = synth(rng, 4000, 10)
    cov[11, 3] = 70000  # outside the 64 Ki-entry context table: materialised fallback with a data-sized table
    prep = dfb.preprocessCoverage(dict(coverage=cov, position=position), groupInfo=group, engine=eng)
    Y = prep["coverageProcessed"].processed()
    assert np.array_equal(Y, O.log2_libm(cov.astype(np.float64) + 32.0))
    F = np.asarray(dfb.calculateStats(prep, models, engine=eng))
    Fo = O.fstats_projector(O.log2_libm(cov.astype(np.float64) + 32.0), models["mod"], models["mod0"])
    ok = np.isfinite(Fo) & (np.abs(Fo) > 1e-8)
    assert np.allclose(F[ok], Fo[ok], rtol=F_RTOL)


# ---------------------------------------------------------------------------------------------
# mergeResults(): fullRegions / fullNullSummary shaped like the reference's .Rdata objects (R/mergeResults.R:199-328)


def test_merge_results_objects(dfb, eng, tmp_path):
    rng = np.random.default_rng(4711)
    n, B = 12, 30
    _, _, _, models = synth(rng, 2000, n)
    perm = random_perms(rng, B, n)
    cut = dfb.calcFstatCutoff("theoretical", 0.05, models=models)
    results, fstats = {}, {}
    for c, N in (("chr1", 9000), ("chr2", 3000), ("chrX", 5000), ("chrY", 1500)):
        cov, position, group, _ = synth(rng, N, n)
        prep = dfb.preprocessCoverage(dict(coverage=cov, position=position), groupInfo=group, engine=eng)
        F = dfb.calculateStats(prep, models, engine=eng)
        fstats[c] = np.asarray(F)
        results[c] = dfb.calculatePvalues(prep, models, F, permIndex=perm, cutoff=1e12 if c == "chr2" else cut, engine=eng)
    assert results["chr2"]["regions"] is None
    opts = dict(cutoffType="theoretical", cutoffFstat=0.05, models=models, nPermute=B)
    got = dfb.mergeResults(results=results, optionsStats=opts, fullFstats=fstats, prefix=str(tmp_path), engine=eng)
    want = O.merge_results(results, B, cutoffFstatUsed=cut)
    assert got["optionsMerge"]["cutoffFstatUsed"] == cut and got["optionsMerge"]["chrs"] == list(results)
    fn, wn = got["fullNullSummary"], want["fullNullSummary"]
    for k in ("stat", "width", "permutation", "area"):
        assert np.array_equal(fn[k], wn[k]), k
    assert list(fn["chr"]) == list(wn["chr"])
    assert np.all(np.diff(fn["area"]) <= 0)
    fr, wr = got["fullRegions"], want["fullRegions"]
    for k in ("start", "end", "area", "value", "indexStart", "indexEnd", "cluster", "fwer", "pvalues"):
        assert np.array_equal(fr[k], wr[k]), k
    assert list(fr["seqnames"]) == list(wr["seqnames"])
    assert np.array_equal(fr["significant"], wr["significant"]) and np.array_equal(fr["significantFWER"], wr["significantFWER"])
    assert np.all(np.diff(fr["area"]) <= 0) and np.all(np.isnan(fr["qvalues"]))
    # the pooled p-values differ from the per-chromosome ones (more nulls), the saved objects round-trip
    saved = np.load(tmp_path / "fullRegions.npz")
    assert np.array_equal(saved["pvalues"], fr["pvalues"]) and np.array_equal(saved["fwer"], fr["fwer"])
    assert np.array_equal(np.load(tmp_path / "fullNullSummary.npz")["area"], fn["area"])
    # empirical cutoff from the concatenated F-statistics (:241-263)
    got2 = dfb.mergeResults(results=results, optionsStats=dict(opts, cutoffType="empirical", cutoffFstat=0.99),
                            fullFstats=fstats, engine=eng)
    q = O.quantile_type7(np.concatenate([fstats[c] for c in results]), 0.99)
    assert got2["optionsMerge"]["cutoffFstatUsed"] == q
    # no chromosome has regions
    none = dfb.mergeResults(results={"chr2": results["chr2"]}, optionsStats=opts, cutoffFstatUsed=cut, engine=eng)
    assert none["fullRegions"] is None and none["fullNullSummary"] is None


# ---------------------------------------------------------------------------------------------
# second-generation null scan forced for small models (by default it is used from 17 samples on)


@pytest.mark.parametrize("n,extra,N,B,alpha", [(12, 0, 20000, 70, 0.05), (14, 2, 15000, 40, 0.02), (16, 1, 9000, 33, 0.2)])
def test_null2_forced_small_models(dfb, n, extra, N, B, alpha):
    rng = np.random.default_rng(555 + n)
    cov, position, group, models = synth(rng, N, n, p_extra=extra, seg=[N // 2, N - N // 2])
    perm = random_perms(rng, B, n)
    cut = dfb.calcFstatCutoff("theoretical", alpha, models=models)
    e = dfb.Engine(0)
    try:
        e.set_option("null2", 2)
        res, ores = run_both(dfb, e, cov, position, group, models, perm, cut)
        assert_pipeline_equal(res, ores)
        dense, screen, cand, diff = debug_screen(e, dfb.preprocessCoverage(dict(coverage=cov, position=position), engine=e)
                                                 ["coverageProcessed"], models, perm, cut)
        assert diff == 0 and dense == screen
    finally:
        e.close()


# ---------------------------------------------------------------------------------------------
# filterData: the kept mask from run events (integer coverage, no divisor) == the expansion kernel == the oracle


@pytest.mark.parametrize("flt,cutoff,n,length,mean_run", [("mean", 2.5, 9, 300007, 37), ("one", 5, 33, 120011, 11),
                                                          ("mean", 0, 5, 70001, 3), ("one", -1, 4, 50000, 50),
                                                          ("mean", 5, 100, 41000, 45)])
def test_filter_events_equals_expansion(dfb, flt, cutoff, n, length, mean_run):
    from test_gpu_parity import random_rle_cols
    rng = np.random.default_rng(8080 + n)
    cols = random_rle_cols(rng, n, length, mean_run, 0.5)
    o = O.filter_data(cols, cutoff=cutoff, filter=flt, returnMean=True)
    for ev in (1, 0):
        e = dfb.Engine(0)
        try:
            e.set_option("filter_events", ev)
            f = dfb.filterData(cols, cutoff=cutoff, filter=flt, returnMean=True, engine=e)
            assert np.array_equal(f["position"][0], o.position[0]) and np.array_equal(f["position"][1], o.position[1]), ev
            assert np.array_equal(f["coverage"].to_numpy(), o.coverage), ev
            assert np.array_equal(f["meanCoverage"], o.meanCoverage), ev
        finally:
            e.close()


# ---------------------------------------------------------------------------------------------
# 16-bit wire format of integer Rle columns (dfb_cov_from_rle): fits / value overflow / length overflow


@pytest.mark.parametrize("case", ["fits", "value_overflow", "length_overflow", "negative", "off", "pinned", "pinned_ragged"])
def test_rle_upload_wire16(dfb, case):
    from derfinder_b200.api import _upload_filtered
    rng = np.random.default_rng(2718)
    n, runs = 8, 135000  # 1.08 M runs in 8 columns: 16-bit wire format and four pipelined sample groups
    cols, dense = [], []
    for s in range(n):
        lens = rng.integers(1, 40, size=runs).astype(np.int64)
        vals = rng.integers(0, 3000, size=runs).astype(np.int32)
        vals[1:][vals[1:] == vals[:-1]] += 1  # proper Rle: neighbours differ
        cols.append([vals, lens])
    N = int(min(l.sum() for _, l in cols))
    for c in cols:  # trim every column to the common length N
        v, l = c
        cs = np.cumsum(l)
        k = int(np.searchsorted(cs, N))
        l = l[:k + 1].copy()
        l[k] -= cs[k] - N
        c[0], c[1] = v[:k + 1].copy(), l
    if case == "value_overflow":
        cols[3][0][1234] = 70000
    elif case == "negative":
        cols[2][0][77] = -5
    elif case == "length_overflow":  # one long run: shorten the column elsewhere to keep the length
        v, l = cols[1]
        extra = 70000 - int(l[500])
        l[500] = 70000
        cs = np.cumsum(l)
        k = int(np.searchsorted(cs, N))
        cols[1][0], l2 = v[:k + 1].copy(), l[:k + 1].copy()
        l2[k] -= cs[k] - N
        cols[1][1] = l2
        assert extra > 0 and l2.sum() == N
    position = (np.array([True]), np.array([N]))
    e = dfb.Engine(0)
    try:
        if case == "off":
            e.set_option("wire16", 0)
        if case.startswith("pinned"):  # page-locked columns are DMAed where they lie (auxiliary stream, per sample group)
            import ctypes as C2
            import torch
            from derfinder_b200 import _lib as L2
            keep, vals, lens = [], [], []
            for v, l in cols:
                for a, dst in ((v.astype(np.int32), vals), (l.astype(np.int32), lens)):
                    t = torch.empty(a.shape, dtype=torch.int32, pin_memory=True)
                    t.numpy()[...] = a
                    keep.append(t)
                    dst.append(t.numpy())
            if case == "pinned_ragged":
                lens[2][10] += 3  # column 3 no longer sums to N: reported, nothing expanded from it
            vp = (C2.c_void_p * n)(*[v.ctypes.data for v in vals])
            lp = (C2.c_void_p * n)(*[l.ctypes.data for l in lens])
            nr = np.asarray([len(v) for v in vals], dtype=np.int64)
            pl = np.asarray([N], dtype=np.int32)
            h = C2.c_void_p()
            rc = e.lib.dfb_cov_from_rle(e.h, n, L2.DFB_I32, vp, lp, L2.ptr(nr, C2.c_int64), N, L2.ptr(pl, C2.c_int32), 1, 1,
                                        C2.byref(h))
            if case == "pinned_ragged":
                assert rc < 0 and b"column 3" in e.lib.dfb_last_error()
                return
            assert rc == 0, e.lib.dfb_last_error()
            cov = dfb.Coverage(e, h)
        else:
            cov = _upload_filtered(e, [(v, l.astype(np.int32)) for v, l in cols], position)
        for s in range(n):
            assert np.array_equal(cov.column(s), np.repeat(cols[s][0], cols[s][1])), (case, s)
    finally:
        e.close()


# ---------------------------------------------------------------------------------------------
# cluster chain by one block == the multi-kernel chain == the oracle (R/findRegions.R:254-271)


@pytest.mark.parametrize("cgap,rgap,two_sided", [(300, 0, False), (50, 0, True), (3000, 5, False), (0, 0, True)])
def test_cluster_chain_one_block(dfb, cgap, rgap, two_sided):
    from test_gpu_parity import assert_regions_equal
    rng = np.random.default_rng(31337 + cgap)
    N = 120000
    # kept bases in islands with gaps of many sizes: clusters of one and of dozens of regions
    lens, vals = [], []
    tot = 0
    while tot < N:
        on = int(rng.integers(20, 900))
        on = min(on, N - tot)
        vals += [False, True]
        lens += [int(rng.choice([1, 7, 40, 250, 301, 2999, 3001, 20000])), on]
        tot += on
    position = (np.array(vals), np.array(lens))
    F = rng.normal(size=N) * 2.0
    F[rng.random(N) < 0.3] = np.round(F[rng.random(N) < 0.3][0], 1)  # some ties
    cutoff = (-1.5, 1.5) if two_sided else 1.0
    want = O.find_regions(position, F, cutoff, maxClusterGap=cgap, maxRegionGap=rgap)
    for one in (1, 0):
        e = dfb.Engine(0)
        try:
            e.set_option("cluster_one_block", one)
            got = dfb.findRegions(position=position, fstats=F, chr="chrT", cutoff=cutoff, maxClusterGap=cgap,
                                  maxRegionGap=rgap, engine=e)
            assert_regions_equal(got, want)
        finally:
            e.close()


def test_find_regions_segment_ir(dfb, eng):
    """findRegions(position = NULL, segmentIR = ..., basic = TRUE) (R/findRegions.R:155-170) and segmentIR given
    next to position (the cache calculatePvalues passes, R/calculatePvalues.R:234-244)."""
    rng = np.random.default_rng(515)
    N = 30000
    F = rng.normal(size=N) * 2.0
    position = (np.array([False, True, False, True, False, True]), np.array([100, 12000, 7, 8000, 400, 10000]))
    seg = O.cluster_maker_rle(position, maxGap=0, ranges=True)
    want = O.find_regions(None, F, 1.2, segmentIR=seg, basic=True)
    got = dfb.findRegions(position=None, fstats=F, cutoff=1.2, segmentIR=seg, basic=True, engine=eng)
    for k in ("area", "width", "stat"):
        assert np.array_equal(got[k], want[k]), k
    full = dfb.findRegions(position=position, fstats=F, chr="chrT", cutoff=1.2, segmentIR=seg, engine=eng)
    ofull = O.find_regions(position, F, 1.2, segmentIR=seg)
    from test_gpu_parity import assert_regions_equal
    assert_regions_equal(full, ofull)
    with pytest.raises(ValueError):
        dfb.findRegions(position=None, fstats=F, cutoff=1.2, segmentIR=seg, engine=eng)  # not basic: position needed
    with pytest.raises(ValueError):
        dfb.findRegions(position=position, fstats=F, cutoff=1.2, segmentIR=(seg[0][:-1], seg[1][:-1]), engine=eng)
