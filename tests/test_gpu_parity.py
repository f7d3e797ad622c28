"""GPU parity tests: the CUDA engine, called through the C ABI (via derfinder_b200.api), against the
oracle and the reference's goldens.  Mirrors tests/testthat/test_Fstats.R, test_findRegions.R,
test-maxRegionGap.R, test_ER_level.R and the filterData cases of test_data.R.

Bars: integer / index work (positions, regions, widths, clusters, permutation ids, p-value counts)
bit-exact; F-statistics bit-exact against the C oracle of the same arithmetic and <= 1e-6 relative
(north_star tolerance) against the reference's projector definition / golden genomeFstats.
"""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import c_oracle  # noqa: E402
from oracle import derfinder_oracle as O  # noqa: E402

F_RTOL = 1e-6  # north_star: F-stats within 1e-6 relative of the reference's double-precision result


@pytest.fixture(scope="module")
def dfb():
    import derfinder_b200 as m
    return m


@pytest.fixture(scope="module")
def eng(dfb):
    e = dfb.Engine(0)
    yield e
    e.close()


def engine_basis(eng, mod, mod0):
    from derfinder_b200 import _lib as L
    mod, mod0 = np.asfortranarray(mod, dtype=np.float64), np.asfortranarray(mod0, dtype=np.float64)
    n, p = mod.shape
    q = np.zeros((p, n))
    cen = C.c_int()
    L.check(eng.lib.dfb_model_basis(mod.ctypes.data_as(L.c_f64p), n, p, mod0.ctypes.data_as(L.c_f64p), mod0.shape[1],
                                    q.ctypes.data_as(L.c_f64p), C.byref(cen)))
    return q.T.copy(), bool(cen.value)


def oracle_idx_fn(eng, mod, mod0, adjustF=0.0):
    Q, cen = engine_basis(eng, mod, mod0)
    return lambda Y, idx: c_oracle.fstats_basis(Y, Q, mod0.shape[1], idx, cen, adjustF)


def synth(rng, N, n, p_extra=0, seg=None, effect=True):
    """Integer coverage with island structure, a 2-group design with `p_extra` adjustment covariates."""
    lam = np.repeat(rng.gamma(2.0, 8.0, size=N // 40 + 1), 40)[:N]
    depth_f = rng.uniform(0.5, 1.5, size=n)
    cov = rng.poisson(lam[:, None] * depth_f[None, :]).astype(np.int32)
    group = np.array(["g1"] * (n // 2) + ["g2"] * (n - n // 2))
    if effect:
        a = N // 5
        cov[a:a + N // 20, : n // 2] *= 3
    depth = np.log2(cov.sum(axis=0) + 32.0)
    cols = [np.ones(n), (group == "g2") * 1.0, depth] + [rng.normal(size=n) for _ in range(p_extra)]
    mod = np.column_stack(cols)
    mod0 = mod[:, [0] + list(range(2, mod.shape[1]))]
    if seg is None:
        seg = [N]
    vals, lens = [], []
    for i, s in enumerate(seg):
        vals += [False, True]
        lens += [int(rng.integers(1, 500)) if i else 1000, s]
    position = (np.array(vals), np.array(lens))
    return cov, position, group, dict(mod=mod, mod0=mod0)


def random_perms(rng, B, n):
    return np.stack([rng.permutation(n) for _ in range(B)]).astype(np.int32)


def assert_regions_equal(a, b, keys=("start", "end", "indexStart", "indexEnd", "cluster", "area", "value")):
    assert (a is None) == (b is None)
    if a is None:
        return
    for k in keys:
        assert np.array_equal(a[k], b[k], equal_nan=True), k
    assert np.array_equal(np.asarray(a["clusterL"], dtype=float), np.asarray(b["clusterL"], dtype=float),
                          equal_nan=True)
    assert list(a["name"]) == list(b["name"])


# ---------------------------------------------------------------------------------------------
# C1: the reference's bundled chr21 slice (tests/testthat/test_Fstats.R:43-87)


def test_golden_preprocess_and_fstats(dfb, eng, golden):
    prep = dfb.preprocessCoverage(dict(coverage=golden.cov, position=golden.position), groupInfo=golden.pop,
                                  scalefac=32, chunksize=1e3, engine=eng)
    assert prep["mclapplyIndex"] == [1000, 434]  # test_findRegions.R:38
    cov = prep["coverageProcessed"]
    assert np.array_equal(cov.processed(), O.log2_libm(golden.cov + 32.0))  # test_findRegions.R:35-37
    assert np.array_equal(prep["meanCoverage"], golden["chr21_prep_meanCoverage"])
    for k in ("CEU", "YRI"):
        assert np.array_equal(prep["groupMeans"][k], golden[f"chr21_prep_groupMeans_{k}"])
    pv, pl = prep["position"]
    assert np.array_equal(pv, golden.position[0]) and np.array_equal(pl, golden.position[1])
    models = dict(mod=golden.mod, mod0=golden.mod0)
    F = dfb.calculateStats(prep, models, engine=eng)
    assert np.max(np.abs(F - golden.fstats) / np.abs(golden.fstats)) < F_RTOL  # golden genomeFstats
    assert np.max(np.abs(F - golden.fstats_matrix) / np.abs(golden.fstats)) < F_RTOL
    fn = oracle_idx_fn(eng, golden.mod, golden.mod0)
    assert np.array_equal(F, fn(O.log2_libm(golden.cov + 32.0), None))  # bit-exact vs the C oracle
    with pytest.raises(ValueError):
        dfb.calculateStats(prep, dict(mod=golden.mod[:30], mod0=golden.mod0[:30]), engine=eng)


def test_golden_calculate_pvalues(dfb, eng, golden):
    from derfinder_b200.rrng import permutation_indices
    prep = dfb.preprocessCoverage(dict(coverage=golden.cov, position=golden.position), groupInfo=golden.pop,
                                  scalefac=32, chunksize=1e3, engine=eng)
    models = dict(mod=golden.mod, mod0=golden.mod0)
    res = dfb.calculatePvalues(prep, models, golden.fstats, nPermute=10, seeds=list(range(1, 11)), chr="chr21",
                               cutoff=1, sampleKind="Rounding", engine=eng)
    r = res["regions"]
    G = lambda k: golden["genomeRegions_" + k]
    assert len(r["start"]) == 33
    for k in ("start", "indexStart", "indexEnd", "cluster", "clusterL", "area", "value", "meanCoverage", "meanCEU",
              "meanYRI"):
        assert np.array_equal(r[k], G(k)), k
    assert np.array_equal(r["end"] - r["start"] + 1, G("width"))
    assert np.array_equal(r["log2FoldChangeYRIvsCEU"], G("log2FoldChangeYRIvsCEU"), equal_nan=True)
    assert list(r["name"]) == list(G("names"))
    assert np.array_equal(res["nullWidths"], G("nullWidths"))
    assert np.array_equal(res["nullPermutation"], G("nullPermutation"))
    assert np.max(np.abs(res["nullStats"] - G("nullStats")) / G("nullStats")) < F_RTOL
    # p-values: exact except inside the declared near-tie band
    nulls = res["nullAreas"]
    for a, x, y in zip(r["area"], r["pvalues"], G("pvalues")):
        if x != y:
            near = np.sum(np.abs(nulls - a) <= 1e-9 * a)
            assert near > 0 and abs(round(x * 541) - round(y * 541)) <= near
    assert np.sum(r["pvalues"] == G("pvalues")) >= 31
    # and bit-exact against the oracle driven with the engine's own arithmetic
    oprep = O.preprocess_coverage(golden.cov, golden.position, groupInfo=golden.pop, chunksize=1e3)
    perm = permutation_indices(range(1, 11), 31, "Rounding")
    ores = O.calculate_pvalues(oprep, golden.mod, golden.mod0, golden.fstats, perm, 1,
                               fstat_idx_fn=oracle_idx_fn(eng, golden.mod, golden.mod0))
    assert np.array_equal(res["nullStats"], ores["nullStats"])
    assert np.array_equal(res["nullAreas"], ores["nullAreas"])
    assert np.array_equal(r["pvalues"], ores["regions"]["pvalues"])
    # default-argument identity (test_Fstats.R:88)
    a = dfb.calculatePvalues(prep, models, golden.fstats, nPermute=1, seeds=[11], chr="chr21", cutoff=1, engine=eng)
    b = dfb.calculatePvalues(prep, models, golden.fstats, nPermute=1, seeds=[11], chr="chr21", cutoff=1,
                             maxRegionGap=0, maxClusterGap=300, adjustF=0, engine=eng)
    assert np.array_equal(a["regions"]["pvalues"], b["regions"]["pvalues"])
    assert np.array_equal(a["nullStats"], b["nullStats"])


def test_golden_bundled_chr21(dfb, eng, golden):
    res = dfb.analyzeChr("chr21", dict(coverage=golden.cov, position=golden.position),
                         dict(mod=golden.mod, mod0=golden.mod0), cutoffFstat=1, cutoffType="manual", nPermute=1,
                         seeds=[20140330], groupInfo=golden.pop, sampleKind="Rounding", engine=eng)
    assert np.max(np.abs(res["fstats"] - golden.fstats_matrix) / golden.fstats) < F_RTOL
    r = res["regions"]["regions"]
    for k in ("start", "indexStart", "indexEnd", "cluster", "clusterL"):
        assert np.array_equal(r[k], golden["chr21_regions_" + k]), k
    assert np.allclose(r["area"], golden["chr21_regions_area"], rtol=F_RTOL)
    assert np.array_equal(res["regions"]["nullWidths"], golden["chr21_regions_nullWidths"])
    assert np.array_equal(r["pvalues"], golden["chr21_regions_pvalues"])


# ---------------------------------------------------------------------------------------------
# literal known answers (test_Fstats.R:72-90, test_findRegions.R:5-18)


def test_known_answers(dfb, eng, golden):
    assert np.array_equal(dfb.calcPval([1, 2], [1, 2, 3, 4], engine=eng), [0.8, 0.6])
    assert np.array_equal(dfb.calculateFWER(10, [1, 2, 3, 4], [1, 1, 3, 3], 3, [1, 2], engine=eng), [1.0, 0.75])
    assert dfb.calcFstatCutoff("empirical", 0.5, fstats=np.repeat([1.0, 2.0], 10), engine=eng) == 1.5
    assert dfb.calcFstatCutoff("manual", 1) == 1
    x = np.repeat([-1.0, 0, 1, 2], 1000)
    s = dfb.getSegmentsRle(x, 0.5, engine=eng)
    assert (list(s["upIndex"][0]), list(s["upIndex"][1])) == ([2001], [4000])
    assert (list(s["dnIndex"][0]), list(s["dnIndex"][1])) == ([1], [1000])
    s = dfb.getSegmentsRle(x + 3, [1, 1], engine=eng)
    assert (list(s["upIndex"][0]), list(s["upIndex"][1])) == ([1], [4000]) and len(s["dnIndex"][0]) == 0


def test_find_regions_known_answers(dfb, eng, golden):
    F = golden.fstats
    small_pos = (np.array([False, True, False, True]), np.array([47407536, 32, 1256, 36]))
    rb = dfb.findRegions(small_pos, F[:68], "chr21", cutoff=1, basic=True, engine=eng)
    ob = O.find_regions(small_pos, F[:68], 1, basic=True)
    assert list(rb["width"]) == [6, 15, 36]
    for k in ("area", "width", "stat"):
        assert np.array_equal(rb[k], ob[k])
    full = dfb.findRegions(golden.position, F, "chr21", engine=eng)  # default cutoff quantile(0.99)
    assert_regions_equal(full, O.find_regions(golden.position, F, O.quantile_type7(F, 0.99)))
    one = dfb.findRegions(golden.position, F, "chr21", cutoff=0, maxRegionGap=100000, maxClusterGap=1000000,
                          engine=eng)
    up = one["name"] == "up"
    assert list(one["start"][up]) == [47407537] and list(one["end"][up]) == [47418068]
    assert_regions_equal(one, O.find_regions(golden.position, F, 0, maxRegionGap=100000, maxClusterGap=1000000))
    with pytest.warns(UserWarning):
        assert dfb.findRegions((np.array([False]), np.array([1])), F, engine=eng) is None
    assert dfb.findRegions(golden.position, F, cutoff=1e9, engine=eng) is None
    for gap, cgap in ((0, 300), (50, 300), (1300, 3000), (10, 5)):
        got = dfb.findRegions(golden.position, F, cutoff=1, maxRegionGap=gap, maxClusterGap=cgap, engine=eng)
        assert_regions_equal(got, O.find_regions(golden.position, F, 1, maxRegionGap=gap, maxClusterGap=cgap))


# ---------------------------------------------------------------------------------------------
# F-statistics across model shapes (bit-exact vs C oracle, 1e-6 vs the projector definition)


@pytest.mark.parametrize("n,extra,N", [(12, 0, 4133), (31, 1, 2999), (60, 3, 1500), (100, 3, 700), (20, 5, 1027),
                                       (9, 0, 513), (48, 2, 2048), (130, 3, 300)])
def test_fstats_shapes(dfb, eng, n, extra, N):
    rng = np.random.default_rng(100 + n)
    cov, position, group, models = synth(rng, N, n, p_extra=extra)
    prep = dfb.preprocessCoverage(dict(coverage=cov, position=position), engine=eng)
    F = dfb.calculateStats(prep, models, adjustF=0.0, engine=eng)
    Y = O.log2_libm(cov + 32.0)
    assert np.array_equal(prep["coverageProcessed"].processed(), Y)
    assert np.array_equal(F, oracle_idx_fn(eng, models["mod"], models["mod0"])(Y, None), equal_nan=True)
    Fo = O.fstats_projector(Y, models["mod"], models["mod0"])
    ok = np.isfinite(Fo) & (Fo > 1e-3)
    assert np.max(np.abs(F[ok] - Fo[ok]) / Fo[ok]) < F_RTOL
    # mixed abs/rel tolerance for vanishing F
    yy = (Y * Y).sum(axis=1)
    assert np.all(np.abs(F - Fo)[np.isfinite(Fo)] <= F_RTOL * np.abs(Fo[np.isfinite(Fo)]) + 1e-12 * yy[np.isfinite(Fo)])
    F2 = dfb.calculateStats(prep, models, adjustF=0.5, engine=eng)
    assert np.array_equal(F2, oracle_idx_fn(eng, models["mod"], models["mod0"], 0.5)(Y, None), equal_nan=True)


def test_fstats_no_intercept_in_null(dfb, eng):
    # makeModels(testIntercept = TRUE): mod0 has no intercept -> centring must be off
    rng = np.random.default_rng(5)
    cov, position, group, models = synth(rng, 1500, 14, p_extra=1)
    mod = models["mod0"]
    mod0 = mod[:, 1:]
    _, cen = engine_basis(eng, mod, mod0)
    assert not cen
    prep = dfb.preprocessCoverage(dict(coverage=cov, position=position), engine=eng)
    F = dfb.calculateStats(prep, dict(mod=mod, mod0=mod0), engine=eng)
    Fo = O.fstats_projector(O.log2_libm(cov + 32.0), mod, mod0)
    ok = Fo > 1e-3
    assert np.max(np.abs(F[ok] - Fo[ok]) / Fo[ok]) < 1e-5  # un-centred identity: documented lower accuracy


def test_model_errors(dfb, eng):
    rng = np.random.default_rng(1)
    cov, position, group, models = synth(rng, 600, 10)
    prep = dfb.preprocessCoverage(dict(coverage=cov, position=position), engine=eng)
    bad = models["mod"].copy()
    bad[:, 2] = bad[:, 0]  # rank deficient
    with pytest.raises(RuntimeError):
        dfb.calculateStats(prep, dict(mod=bad, mod0=models["mod0"]), engine=eng)
    with pytest.raises(RuntimeError):  # mod0 not nested
        dfb.calculateStats(prep, dict(mod=models["mod"], mod0=rng.normal(size=(10, 2))), engine=eng)


# ---------------------------------------------------------------------------------------------
# permutation nulls + regions + p-values vs the oracle (bit-exact on the engine's arithmetic)


def run_both(dfb, eng, cov, position, group, models, perm, cutoff, gap=0, cgap=300, adjustF=0.0):
    prep = dfb.preprocessCoverage(dict(coverage=cov, position=position), groupInfo=group, engine=eng)
    F = dfb.calculateStats(prep, models, adjustF=adjustF, engine=eng)
    res = dfb.calculatePvalues(prep, models, F, permIndex=perm, cutoff=cutoff, maxRegionGap=gap, maxClusterGap=cgap,
                               adjustF=adjustF, engine=eng)
    oprep = O.preprocess_coverage(cov, position, groupInfo=group)
    ores = O.calculate_pvalues(oprep, models["mod"], models["mod0"], F, perm, cutoff, maxRegionGap=gap,
                               maxClusterGap=cgap, adjustF=adjustF,
                               fstat_idx_fn=oracle_idx_fn(eng, models["mod"], models["mod0"], adjustF))
    return res, ores


def assert_pipeline_equal(res, ores):
    if ores["regions"] is None:
        assert res["regions"] is None
        return
    assert_regions_equal(res["regions"], ores["regions"])
    for k in ores["regions"]:
        if k.startswith("mean") or k.startswith("log2FoldChange"):
            assert np.array_equal(res["regions"][k], ores["regions"][k], equal_nan=True), k
    assert np.array_equal(res["nullWidths"], ores["nullWidths"])
    assert np.array_equal(res["nullPermutation"], ores["nullPermutation"])
    assert np.array_equal(res["nullStats"], ores["nullStats"])
    assert np.array_equal(res["nullAreas"], ores["nullAreas"])
    assert np.array_equal(res["regions"]["pvalues"], ores["regions"]["pvalues"], equal_nan=True)


@pytest.mark.parametrize("n,extra,N,B,seg", [(12, 0, 20000, 12, [7000, 33, 5000, 7967]),
                                             (31, 1, 6000, 7, [6000]),
                                             (60, 3, 3000, 5, [1, 31, 32, 33, 2903]),
                                             (10, 0, 70001, 9, [20000, 50001]),
                                             (24, 3, 5000, 100, [5000]),    # wide chunks: 48 + 48 + 4 permutations
                                             (20, 5, 4000, 70, [4000])])    # p = 8: 32-permutation chunks of 224 columns
def test_pipeline_vs_oracle(dfb, eng, n, extra, N, B, seg):
    rng = np.random.default_rng(n * 7 + B)
    cov, position, group, models = synth(rng, N, n, p_extra=extra, seg=seg)
    perm = random_perms(rng, B, n)
    cut = dfb.calcFstatCutoff("theoretical", 0.05, models=models)
    res, ores = run_both(dfb, eng, cov, position, group, models, perm, cut)
    assert len(ores["nullStats"]) > 0
    assert_pipeline_equal(res, ores)


def test_pipeline_region_gap_and_two_sided(dfb, eng):
    rng = np.random.default_rng(77)
    cov, position, group, models = synth(rng, 9000, 12, seg=[3000, 200, 40, 5760])
    perm = random_perms(rng, 6, 12)
    # maxRegionGap > 0: regions may bridge filtered-out positions (cluster quirk of SURVEY A.4)
    res, ores = run_both(dfb, eng, cov, position, group, models, perm, 3.0, gap=600, cgap=700)
    assert_pipeline_equal(res, ores)
    # cutoff pair with L >= 0: the "dn" side (F <= L) is live for observed and null regions
    res, ores = run_both(dfb, eng, cov, position, group, models, perm, [0.05, 4.0])
    assert np.any(res["regions"]["name"] == "dn")
    assert_pipeline_equal(res, ores)
    # adjustF in the denominator (R/calculateStats.R:24-26)
    res, ores = run_both(dfb, eng, cov, position, group, models, perm, 2.0, adjustF=0.25)
    assert_pipeline_equal(res, ores)
    # dense exceedance (53% of bases pass in the golden test): low cutoff
    res, ores = run_both(dfb, eng, cov, position, group, models, perm, 0.3)
    assert_pipeline_equal(res, ores)
    # nothing passes -> list of NULLs (R/calculatePvalues.R:245-251)
    prep = dfb.preprocessCoverage(dict(coverage=cov, position=position), engine=eng)
    F = dfb.calculateStats(prep, models, engine=eng)
    out = dfb.calculatePvalues(prep, models, F, permIndex=perm, cutoff=1e12, engine=eng)
    assert out["regions"] is None and out["nullStats"] is None


def test_exceedance_overflow_and_batching(dfb, eng):
    rng = np.random.default_rng(3)
    cov, position, group, models = synth(rng, 30000, 12, seg=[30000])
    perm = random_perms(rng, 11, 12)
    base, ores = run_both(dfb, eng, cov, position, group, models, perm, 1.0)
    assert_pipeline_equal(base, ores)
    e2 = dfb.Engine(0)
    try:
        e2.set_option("null_capacity_mb", 0.01)  # forces the count-then-refill pass
        e2.set_option("perm_batch", 3)           # 4 launches for 11 permutations
        prep = dfb.preprocessCoverage(dict(coverage=cov, position=position), groupInfo=group, engine=e2)
        F = dfb.calculateStats(prep, models, engine=e2)
        res = dfb.calculatePvalues(prep, models, F, permIndex=perm, cutoff=1.0, engine=e2)
        assert_pipeline_equal(res, ores)
    finally:
        e2.close()


def test_no_nulls_gives_na(dfb, eng):
    rng = np.random.default_rng(9)
    cov, position, group, models = synth(rng, 4000, 10)
    prep = dfb.preprocessCoverage(dict(coverage=cov, position=position), engine=eng)
    F = dfb.calculateStats(prep, models, engine=eng)
    cut = float(np.nanmax(F)) * 0.999
    perm = np.arange(10, dtype=np.int32)[None, ::-1].copy()
    res = dfb.calculatePvalues(prep, models, F, permIndex=perm, cutoff=cut, engine=eng)
    if len(res["nullStats"]) == 0:
        assert np.all(np.isnan(res["regions"]["pvalues"]))


# ---------------------------------------------------------------------------------------------
# p-value / FWER / quantile on random inputs incl. ties


def test_pval_fwer_quantile_random(dfb, eng):
    rng = np.random.default_rng(11)
    nulls = np.round(rng.gamma(2.0, 3.0, size=50000), 2)  # many exact ties
    areas = np.concatenate((np.round(rng.gamma(2.0, 3.0, size=3000), 2), [0.0, 1e9, nulls.min(), nulls.max()]))
    assert np.array_equal(dfb.calcPval(areas, nulls, engine=eng), O.calc_pval(areas, nulls))
    perm = rng.integers(1, 41, size=nulls.size).astype(np.int32)
    perm[perm == 7] = 8  # an empty permutation -> cutoffFstatUsed
    assert np.array_equal(dfb.calculateFWER(2.5, nulls, perm, 40, areas, engine=eng),
                          O.calculate_fwer(2.5, nulls, perm, 40, areas))
    x = rng.normal(size=100001)
    x[::97] = np.nan
    for p in (0.0, 0.5, 0.99, 1.0, 0.123456):
        assert dfb.calcFstatCutoff("empirical", p, fstats=x, engine=eng) == O.quantile_type7(x, p)


# ---------------------------------------------------------------------------------------------
# filterData (test_data.R:174-191) and random Rle inputs


def test_filter_data_golden(dfb, eng, golden):
    f = dfb.filterData(golden.raw_rle, cutoff=0.5, filter="mean", engine=eng)
    assert f["coverage"].shape == (273, 31)
    f = dfb.filterData(golden.raw_rle, cutoff=10, engine=eng)
    assert list(f["position"][0]) == [False] and list(f["position"][1]) == [48129895]
    f = dfb.filterData(golden.raw_rle, cutoff=0, returnMean=True, engine=eng)
    assert np.array_equal(f["coverage"].to_numpy(), golden.cov)
    assert np.array_equal(f["position"][0], golden.position[0]) and np.array_equal(f["position"][1], golden.position[1])
    assert np.array_equal(f["meanCoverage"], golden["chr21_prep_meanCoverage"])
    f = dfb.filterData(golden.raw_rle, cutoff=-1, returnCoverage=False, engine=eng)
    assert list(f["position"][0]) == [True] and list(f["position"][1]) == [48129895]
    with pytest.raises(ValueError):
        dfb.filterData(golden.raw_rle, filter="two", engine=eng)
    # filt2 <- filterData(filt1$coverage[, 1:2], 5, index = filt1$position) (R/filterData.R:85-88)
    f1 = dfb.filterData(golden.raw_rle, cutoff=0, engine=eng)
    d1 = f1["coverage"].to_numpy()
    f2 = dfb.filterData(d1[:, :2], cutoff=1, index=f1["position"], returnMean=True, engine=eng)
    cols2 = [O.rle_encode(d1[:, s]) for s in range(2)]
    o2 = O.filter_data(cols2, cutoff=1, index=f1["position"], returnMean=True)
    assert np.array_equal(f2["coverage"].to_numpy(), o2.coverage)
    assert np.array_equal(f2["position"][0], o2.position[0]) and np.array_equal(f2["position"][1], o2.position[1])
    assert np.array_equal(f2["meanCoverage"], o2.meanCoverage)
    # preprocessCoverage(colsubset = 1:2, cutoff = 1, groupInfo = factor(letters[1:2])) (test_findRegions.R:25-44)
    ps = dfb.preprocessCoverage(dict(coverage=golden.cov, position=golden.position), groupInfo=np.array(["a", "b"]),
                                cutoff=1, colsubset=[1, 2], chunksize=1e3, engine=eng)
    ds = O.filter_data([O.rle_encode(golden.cov[:, s]) for s in range(2)], 1, golden.position)
    assert np.array_equal(ps["position"][1], ds.position[1])
    assert np.array_equal(ps["coverageProcessed"].processed(), O.log2_libm(ds.coverage + 32.0))
    assert np.array_equal(ps["groupMeans"]["a"], ds.coverage[:, 0].astype(float))
    assert np.array_equal(ps["groupMeans"]["b"], ds.coverage[:, 1].astype(float))


def random_rle_cols(rng, n, length, mean_run, zero_frac, as_float=False):
    cols = []
    for _ in range(n):
        lens = []
        total = 0
        while total < length:
            l = int(rng.geometric(1.0 / mean_run))
            l = min(l, length - total)
            lens.append(l)
            total += l
        lens = np.asarray(lens, dtype=np.int64)
        vals = rng.poisson(6.0, size=len(lens))
        vals[rng.random(len(lens)) < zero_frac] = 0
        # long empty stretches like a real chromosome
        k = len(lens) // 3
        vals[k:k + max(1, len(lens) // 4)] = 0
        v, l = O.rle_normalize(vals.astype(np.float64) * (1.25 if as_float else 1), lens)
        cols.append((v if as_float else v.astype(np.int32), l))
    return cols


@pytest.mark.parametrize("kind,cutoff,flt,norm", [("int", 5, "one", False), ("int", 2.5, "mean", False),
                                                  ("float", 4.0, "one", False), ("int", 3.0, "mean", True),
                                                  ("int", None, "one", False), ("int", 7, "one", True)])
def test_filter_data_random(dfb, eng, kind, cutoff, flt, norm):
    import zlib
    rng = np.random.default_rng(zlib.crc32(repr((kind, cutoff, flt, norm)).encode()))
    n, length = 7, 200003
    cols = random_rle_cols(rng, n, length, 37, 0.5, as_float=(kind == "float"))
    tm = rng.uniform(30e6, 120e6, size=n) if norm else None
    f = dfb.filterData(cols, cutoff=cutoff, filter=flt, totalMapped=tm, targetSize=80e6, returnMean=True, engine=eng)
    o = O.filter_data(cols, cutoff=cutoff, filter=flt, totalMapped=tm, targetSize=80e6, returnMean=True)
    assert np.array_equal(f["coverage"].to_numpy(), o.coverage)
    assert np.array_equal(f["meanCoverage"], o.meanCoverage)
    if cutoff is None:
        assert f["position"] is None
    else:
        assert np.array_equal(f["position"][0], o.position[0]) and np.array_equal(f["position"][1], o.position[1])


# ---------------------------------------------------------------------------------------------
# regionMatrix (test_ER_level.R:5-70)


def test_region_matrix(dfb, eng):
    rng = np.random.default_rng(20160606)
    x, y, z = (np.round(rng.uniform(0, 10, size=10000)) for _ in range(3))
    dense = np.stack([x, y, z], axis=1)
    lib = dense.sum(axis=0)
    cols = [O.rle_encode(dense[:, s]) for s in range(3)]
    rm = dfb.regionMatrix({"chr21": cols}, maxRegionGap=10, maxClusterGap=300, L=36, totalMapped=lib, targetSize=4e4,
                          engine=eng)["chr21"]
    o = O.region_matrix(cols, cutoff=5, L=36, totalMapped=lib, targetSize=4e4, maxRegionGap=10, maxClusterGap=300)
    assert_regions_equal(rm["regions"], o["regions"])
    assert np.allclose(rm["coverageMatrix"], o["coverageMatrix"], rtol=1e-12, atol=0)
    assert len(rm["bpCoverage"]) == len(o["regions"]["start"])
    assert rm["bpCoverage"][0].shape == (int(o["regions"]["indexEnd"][0] - o["regions"]["indexStart"][0] + 1), 3)
    # filter first, then runFilter = FALSE (regionMat2), list-with-position inputs (regionMat4/5)
    filt = dfb.filterData(cols, returnMean=True, filter="mean", cutoff=5, totalMapped=lib, targetSize=4e4, engine=eng)
    rm2 = dfb.regionMatrix({"chr21": filt}, maxRegionGap=10, maxClusterGap=300, L=36, runFilter=False, engine=eng)["chr21"]
    assert_regions_equal(rm2["regions"], rm["regions"])
    assert np.array_equal(rm2["coverageMatrix"], rm["coverageMatrix"])
    rm5 = dfb.regionMatrix({"chr21": dict(coverage=cols, position=(np.array([True]), np.array([10000])))},
                           maxRegionGap=10, maxClusterGap=300, L=36, totalMapped=lib, targetSize=4e4, engine=eng)["chr21"]
    assert np.array_equal(rm5["coverageMatrix"], rm["coverageMatrix"])
    rm3 = dfb.regionMatrix({"chr21": cols}, maxRegionGap=10, maxClusterGap=30000, L=36, totalMapped=lib,
                           targetSize=4e4, engine=eng)["chr21"]
    assert np.all(rm3["regions"]["cluster"] == 1)
    # integer coverage, defaults (totalMapped == targetSize): exact; vector L hits only the last column
    icols = [(v.astype(np.int32), l) for v, l in cols]
    rmi = dfb.regionMatrix({"chr21": icols}, cutoff=5, L=[36, 50, 76], engine=eng)["chr21"]
    oi = O.region_matrix(icols, cutoff=5, L=[36, 50, 76])
    assert np.array_equal(rmi["coverageMatrix"], oi["coverageMatrix"])
    none = dfb.regionMatrix({"chr21": icols}, cutoff=50, L=36, engine=eng)["chr21"]
    assert none["regions"] is None and none["coverageMatrix"] is None


# ---------------------------------------------------------------------------------------------
# tcgen05 screening contraction: candidate mask must contain every exceedance; screen + exact
# refinement must reproduce the dense FP64 kernel bit for bit


def debug_screen(eng, cov_handle, models, perm, cutoff, adjustF=0.0):
    from derfinder_b200 import _lib as L
    mod, mod0 = np.asfortranarray(models["mod"]), np.asfortranarray(models["mod0"])
    perm = np.ascontiguousarray(perm, dtype=np.int32)
    out = [C.c_int64() for _ in range(4)]
    L.check(eng.lib.dfb_debug_screen(eng.h, cov_handle.h, mod.ctypes.data_as(L.c_f64p), mod.shape[1],
                                     mod0.ctypes.data_as(L.c_f64p), mod0.shape[1], float(adjustF),
                                     L.ptr(perm, C.c_int32), perm.shape[0], float(cutoff), *[C.byref(o) for o in out]))
    return tuple(o.value for o in out)  # exact_dense, exact_screen, candidates, mask bits differing


@pytest.mark.parametrize("n,extra,N,B,alpha", [(12, 0, 50000, 37, 0.05), (12, 0, 4099, 100, 0.2),
                                                 (31, 1, 20000, 33, 0.01), (60, 3, 9000, 50, 1e-3),
                                                 (20, 5, 7000, 19, 0.05), (9, 0, 5000, 70, 0.5),
                                                 (80, 3, 3000, 24, 0.01)])
def test_screen_contains_all_exceedances(dfb, eng, n, extra, N, B, alpha):
    rng = np.random.default_rng(1000 + n + B)
    cov, position, group, models = synth(rng, N, n, p_extra=extra)
    prep = dfb.preprocessCoverage(dict(coverage=cov, position=position), engine=eng)
    perm = random_perms(rng, B, n)
    cut = dfb.calcFstatCutoff("theoretical", alpha, models=models)
    dense, screen, cand, diff = debug_screen(eng, prep["coverageProcessed"], models, perm, cut)
    assert diff == 0, f"{diff} mask bits differ between the dense kernel and screen+refine"
    assert dense == screen
    assert cand >= dense
    # the candidate set is tight: error-bound slack only
    assert cand <= dense * 1.05 + 0.001 * N * B + 64, (cand, dense)


def test_screen_off_equals_on(dfb, eng):
    rng = np.random.default_rng(31)
    cov, position, group, models = synth(rng, 25000, 12, seg=[9000, 16000])
    perm = random_perms(rng, 40, 12)
    cut = dfb.calcFstatCutoff("theoretical", 0.05, models=models)
    res_on, ores = run_both(dfb, eng, cov, position, group, models, perm, cut)
    assert_pipeline_equal(res_on, ores)
    e2 = dfb.Engine(0)
    try:
        e2.set_option("screen", 0)
        prep = dfb.preprocessCoverage(dict(coverage=cov, position=position), groupInfo=group, engine=e2)
        F = dfb.calculateStats(prep, models, engine=e2)
        res_off = dfb.calculatePvalues(prep, models, F, permIndex=perm, cutoff=cut, engine=e2)
        assert_pipeline_equal(res_off, ores)
        assert e2.kernel_time(3)[1] == 0 and eng.kernel_time(3)[1] > 0  # tcgen05 path really ran / really off
    finally:
        e2.close()


def test_fused_degenerate_rows(dfb, eng):
    """Bases whose samples all carry the same coverage have yc = 0 (or a constant rounding residue):
    F is 0/0, Inf or garbage in the exact contract; the fused screen (which drops the constant
    column) must still flag exactly what the dense FP64 kernel flags."""
    rng = np.random.default_rng(77)
    N, n = 12000, 12
    cov, position, group, models = synth(rng, N, n)
    flat = rng.random(N) < 0.3
    cov[flat, :] = rng.integers(0, 40, size=flat.sum())[:, None]  # all samples equal at 30% of the bases
    prep = dfb.preprocessCoverage(dict(coverage=cov, position=position), engine=eng)
    perm = random_perms(rng, 45, n)
    for alpha, adj in ((0.05, 0.0), (0.3, 0.0), (0.05, 1e-3)):
        cut = dfb.calcFstatCutoff("theoretical", alpha, models=models)
        dense, screen, cand, diff = debug_screen(eng, prep["coverageProcessed"], models, perm, cut, adjustF=adj)
        assert diff == 0 and dense == screen, (alpha, adj, diff, dense, screen)


def test_fused_off_equals_on(dfb, eng):
    rng = np.random.default_rng(32)
    cov, position, group, models = synth(rng, 25000, 14, p_extra=2, seg=[9000, 16000])
    perm = random_perms(rng, 70, 14)
    cut = dfb.calcFstatCutoff("theoretical", 0.02, models=models)
    res_on, ores = run_both(dfb, eng, cov, position, group, models, perm, cut)
    assert_pipeline_equal(res_on, ores)
    e2 = dfb.Engine(0)
    try:
        e2.set_option("fused", 0)
        prep = dfb.preprocessCoverage(dict(coverage=cov, position=position), groupInfo=group, engine=e2)
        F = dfb.calculateStats(prep, models, engine=e2)
        res_off = dfb.calculatePvalues(prep, models, F, permIndex=perm, cutoff=cut, engine=e2)
        assert_pipeline_equal(res_off, ores)
    finally:
        e2.close()


def test_preprocess_beyond_context_table(dfb, eng):
    """Integer coverage beyond the context's 64 Ki-entry log2 table takes the data-sized table path;
    values must still be bit-identical to libm log2(x + scalefac)."""
    rng = np.random.default_rng(5)
    N, n = 3000, 6
    cov = rng.integers(0, 50, size=(N, n)).astype(np.int32)
    cov[17, 2] = 70000
    cov[2999, 5] = 3_000_000
    position = (np.array([True]), np.array([N]))
    prep = dfb.preprocessCoverage(dict(coverage=cov, position=position), engine=eng)
    oprep = O.preprocess_coverage(cov, position)
    assert np.array_equal(prep["coverageProcessed"].processed(), oprep["coverageProcessed"])
    # and a following ordinary call still uses the context table
    cov2 = rng.integers(0, 50, size=(N, n)).astype(np.int32)
    prep2 = dfb.preprocessCoverage(dict(coverage=cov2, position=position), engine=eng)
    assert np.array_equal(prep2["coverageProcessed"].processed(), O.preprocess_coverage(cov2, position)["coverageProcessed"])


def test_pooled_pvalues_single_rank(dfb, eng):
    """pool.genomewide_pvalues (sorted-list pooling, no second sort) on one rank must reproduce the
    p-values dfb_calculate_pvalues computed itself, and .calculateFWER of the oracle."""
    import torch
    import bench
    from derfinder_b200 import pool
    from derfinder_b200 import _lib as L
    from derfinder_b200.rrng import permutation_indices
    cfg = bench.WORKLOADS["tiny"]
    device = torch.device("cuda", 0)
    work = bench.make_workload(cfg, 11, torch=torch, device=device)
    perm = permutation_indices(range(1, cfg["B"] + 1), cfg["n"], "Rejection")
    cut = bench.theoretical_cutoff(cfg["alpha"], work["mod"], work["mod0"])
    e2 = dfb.Engine(0, stream=torch.cuda.current_stream().cuda_stream)
    try:
        step = bench.Step(e2, work, perm, cut, torch)
        (R, M), rh, nh, cov = step.device_pass(keep=True)
        assert R > 0 and M > 0
        p_dev, fwer, total = pool.genomewide_pvalues(e2, rh, nh, cfg["B"], cut, torch, device)
        torch.cuda.synchronize()
        assert total == M
        pv, order, area = np.empty(R), np.empty(R, dtype=np.int32), np.empty(R)
        L.check(e2.lib.dfb_regions_get_pvalues(rh, pv.ctypes.data_as(L.c_f64p)))
        L.check(e2.lib.dfb_regions_get_order(rh, order.ctypes.data_as(L.c_i32p)))
        L.check(e2.lib.dfb_regions_get(rh, None, None, None, area.ctypes.data_as(L.c_f64p), None, None, None, None))
        assert np.array_equal(p_dev.cpu().numpy()[order], pv)
        na, npm = np.empty(M), np.empty(M, dtype=np.int32)
        L.check(e2.lib.dfb_nulls_get(nh, 0, None, None, na.ctypes.data_as(L.c_f64p), npm.ctypes.data_as(L.c_i32p)))
        want = O.calculate_fwer(cut, na, npm, cfg["B"], area)
        assert np.array_equal(fwer.cpu().numpy()[order], want)
        # a record capacity that is too small is noticed by every rank and the exchange repeated
        for k in list(pool._capacity):
            pool._capacity[k] = 8
        p_again, fwer_again, total_again = pool.genomewide_pvalues(e2, rh, nh, cfg["B"], cut, torch, device)
        torch.cuda.synchronize()
        assert total_again == M and min(pool._capacity.values()) >= R
        assert torch.equal(p_again, p_dev) and torch.equal(fwer_again, fwer)
        # mergeResults pools abs(stat) * width instead of the summed areas (R/mergeResults.R:231)
        p_m, fwer_m, _ = pool.genomewide_pvalues(e2, rh, nh, cfg["B"], cut, torch, device, merge_area=True)
        torch.cuda.synchronize()
        ns, nw = np.empty(M), np.empty(M, dtype=np.int32)
        L.check(e2.lib.dfb_nulls_get(nh, 0, ns.ctypes.data_as(L.c_f64p), nw.ctypes.data_as(L.c_i32p), None, None))
        marea = np.abs(ns) * nw
        assert np.array_equal(p_m.cpu().numpy()[order], O.calc_pval(area, np.repeat(marea, 2)))
        assert np.array_equal(fwer_m.cpu().numpy()[order], O.calculate_fwer(cut, marea, npm, cfg["B"], area))
        e2.lib.dfb_regions_free(rh)
        e2.lib.dfb_nulls_free(nh)
        e2.lib.dfb_cov_free(cov)
    finally:
        e2.close()


# ---------------------------------------------------------------------------------------------
# full-size properties (BASELINE.json configs[1]: N = 4.8 M kept bases, 12 samples, 100 permutations):
# the oracle cannot run at this size in seconds, so the checks are size-independent properties


@pytest.fixture(scope="module")
def full_c2(dfb):
    import torch
    import bench
    from derfinder_b200.rrng import permutation_indices
    cfg = bench.WORKLOADS["c2"]
    device = torch.device("cuda", 0)
    work = bench.make_workload(cfg, 20140923 + 2, torch=torch, device=device)
    perm = permutation_indices(range(1, cfg["B"] + 1), cfg["n"], "Rejection")
    perm[3] = np.arange(cfg["n"])  # one identity relabelling: its nulls must equal the observed regions
    cut = bench.theoretical_cutoff(cfg["alpha"], work["mod"], work["mod0"])
    return cfg, work, perm, cut, torch, bench


def _full_pass(dfb, full, **opts):
    from derfinder_b200 import _lib as L
    cfg, work, perm, cut, torch, bench = full
    e2 = dfb.Engine(0, stream=torch.cuda.current_stream().cuda_stream)
    for k, v in opts.items():
        e2.set_option(k, v)
    try:
        step = bench.Step(e2, work, perm, cut, torch)
        (R, M), rh, nh, cov = step.device_pass(keep=True)
        lib = e2.lib
        out = dict(R=R, M=M)
        out["area"], out["value"], out["pv"] = np.empty(R), np.empty(R), np.empty(R)
        out["is"], out["ie"], out["order"] = (np.empty(R, dtype=np.int32) for _ in range(3))
        L.check(lib.dfb_regions_get(rh, None, None, out["value"].ctypes.data_as(L.c_f64p),
                                    out["area"].ctypes.data_as(L.c_f64p), L.ptr(out["is"], C.c_int32),
                                    L.ptr(out["ie"], C.c_int32), None, None))
        L.check(lib.dfb_regions_get_pvalues(rh, out["pv"].ctypes.data_as(L.c_f64p)))
        L.check(lib.dfb_regions_get_order(rh, L.ptr(out["order"], C.c_int32)))
        out["nstat"], out["narea"] = np.empty(M), np.empty(M)
        out["nw"], out["nperm"] = np.empty(M, dtype=np.int32), np.empty(M, dtype=np.int32)
        L.check(lib.dfb_nulls_get(nh, 0, out["nstat"].ctypes.data_as(L.c_f64p), L.ptr(out["nw"], C.c_int32),
                                  out["narea"].ctypes.data_as(L.c_f64p), L.ptr(out["nperm"], C.c_int32)))
        dup = [np.empty(2 * M), np.empty(2 * M, dtype=np.int32), np.empty(2 * M), np.empty(2 * M, dtype=np.int32)]
        L.check(lib.dfb_nulls_get(nh, 1, dup[0].ctypes.data_as(L.c_f64p), L.ptr(dup[1], C.c_int32),
                                  dup[2].ctypes.data_as(L.c_f64p), L.ptr(dup[3], C.c_int32)))
        out["dup"] = dup
        cnt = np.empty(cfg["B"], dtype=np.int64)
        L.check(lib.dfb_nulls_counts_per_perm(nh, cnt.ctypes.data_as(L.c_i64p)))
        out["per_perm"] = cnt
        lib.dfb_regions_free(rh)
        lib.dfb_nulls_free(nh)
        lib.dfb_cov_free(cov)
        return out
    finally:
        e2.close()


def test_full_size_properties(dfb, full_c2):
    cfg, work, perm, cut, torch, bench = full_c2
    a = _full_pass(dfb, full_c2)
    R, M, B = a["R"], a["M"], cfg["B"]
    assert R > 1000 and M > 100000
    # regions come back ordered by decreasing area (stable), p-values non-decreasing along that order
    assert np.all(np.diff(a["area"]) <= 0)
    assert np.all(np.diff(a["pv"]) >= 0)
    assert a["pv"].min() >= 1.0 / (2 * M + 1) and a["pv"].max() <= 1.0
    # widths / index ranges are consistent, regions do not overlap, means pass the cutoff
    w = a["ie"].astype(np.int64) - a["is"] + 1
    assert w.min() >= 1 and a["ie"].max() <= cfg["N"]
    o = np.argsort(a["is"])
    assert np.all(a["is"][o][1:] > a["ie"][o][:-1])
    assert np.all(a["value"] >= cut) and np.allclose(a["area"], a["value"] * w, rtol=1e-12)
    # nulls: ordered by permutation, counts add up, every null region's mean reaches the cutoff
    assert np.all(np.diff(a["nperm"]) >= 0) and a["nperm"].min() >= 1 and a["nperm"].max() <= B
    assert np.array_equal(np.bincount(a["nperm"], minlength=B + 1)[1:], a["per_perm"]) and a["per_perm"].sum() == M
    assert np.all(a["nstat"] >= cut) and np.allclose(a["narea"], a["nstat"] * a["nw"], rtol=1e-12)
    # the identity relabelling (permutation 4) reproduces the observed regions bit for bit
    sel = a["nperm"] == 4
    inv = np.empty(R, dtype=np.int64)
    inv[a["order"]] = np.arange(R)  # position of up-then-dn region k in the sorted output
    assert sel.sum() == R
    assert np.array_equal(a["narea"][sel], a["area"][inv]) and np.array_equal(a["nstat"][sel], a["value"][inv])
    assert np.array_equal(a["nw"][sel], w[inv])
    # the duplicated layout is [r1..rk, r1..rk] per permutation (R/calculatePvalues.R:346-354)
    pre = np.concatenate(([0], np.cumsum(a["per_perm"])))
    for b in (0, 3, B - 1):
        k = a["per_perm"][b]
        blk = a["dup"][2][2 * pre[b]: 2 * pre[b] + 2 * k]
        assert np.array_equal(blk[:k], a["narea"][pre[b]:pre[b + 1]]) and np.array_equal(blk[k:], blk[:k])
    assert np.array_equal(np.sort(a["dup"][2]), np.sort(np.concatenate([a["narea"], a["narea"]])))
    # p-values against the oracle's .calcPval on the (much smaller) region list
    assert np.array_equal(a["pv"], O.calc_pval(a["area"], a["dup"][2]))


def test_full_size_paths_agree(dfb, full_c2):
    """fused tcgen05 path == two-kernel screen path == dense FP64 path, bit for bit, at full size."""
    ref = _full_pass(dfb, full_c2)
    for opts in (dict(fused=0), dict(screen=0)):
        b = _full_pass(dfb, full_c2, **opts)
        for k in ("area", "value", "pv", "is", "ie", "nstat", "narea", "nw", "nperm"):
            assert np.array_equal(ref[k], b[k]), (opts, k)


@pytest.mark.parametrize("opts", [dict(ys_smem=0), dict(ys_smem=0, epi_warps=2), dict(epi_warps=2), dict(b_stages=3),
                                  dict(ys_smem=1, b_stages=2, epi_warps=4), dict(k2_waves=1), dict(k2_waves=3)])
def test_fused_kernel_variants_agree(dfb, eng, opts):
    """Every layout variant of the fused null kernel (streamed FP64 tile, several epilogue warps per
    TMEM quarter, deeper B staging) must reproduce the oracle pipeline bit for bit on a
    candidate-rich case."""
    rng = np.random.default_rng(404)
    cov, position, group, models = synth(rng, 21000, 16, p_extra=1, seg=[8000, 13000])
    perm = random_perms(rng, 45, 16)
    cut = dfb.calcFstatCutoff("theoretical", 0.1, models=models)
    e2 = dfb.Engine(0)
    try:
        for k, v in opts.items():
            e2.set_option(k, v)
        res, ores = run_both(dfb, e2, cov, position, group, models, perm, cut)
        assert_pipeline_equal(res, ores)
        assert e2.kernel_time(3)[1] > 0  # the tcgen05 path really ran
    finally:
        e2.close()


def test_coverage_to_exon(dfb, eng):
    """coverageToExon (R/coverageToExon.R:206-222): exon x sample sums of unfiltered coverage."""
    rng = np.random.default_rng(99)
    n = 5
    full, exons, dense_by_chr = {}, {}, {}
    for chrom, Lc in (("chr21", 30000), ("chrX", 12345)):
        cols = random_rle_cols(rng, n, Lc, mean_run=40, zero_frac=0.6)
        full[chrom] = cols
        dense_by_chr[chrom] = np.column_stack([O.rle_decode(v, l) for v, l in cols])
        cuts = np.sort(rng.choice(np.arange(1, Lc), size=60, replace=False))
        st, en = cuts[0::2] + 1, cuts[1::2]
        keep = en >= st
        exons[chrom] = (st[keep], en[keep])
    for Lval in (76, np.arange(1, n + 1) * 10.0):
        got = dfb.coverageToExon(full, exons, L_=Lval, engine=eng)
        for chrom in full:
            want = O.coverage_to_exon(dense_by_chr[chrom], exons[chrom][0], exons[chrom][1], L=Lval)
            assert np.array_equal(got[chrom], want), chrom


@pytest.mark.gpu
def test_rail_matrix(dfb, eng):
    """railMatrix (R/railMatrix.R:176-290): ERs of the mean track, then per-sample Rle view sums of
    the library-size normalised coverage -- bit for bit against the oracle, for scalar and
    per-sample L, independent of the region grouping (tests/testthat/test_ER_level.R:156), and the
    same regions as regionMatrix on the same coverage (:158)."""
    rng = np.random.default_rng(2718)
    n, Lc = 6, 40000
    cols = random_rle_cols(rng, n, Lc, mean_run=45, zero_frac=0.5)
    dense = np.column_stack([O.rle_decode(v, l) for v, l in cols]).astype(np.float64)
    acc = np.zeros(Lc)
    for s in range(n):
        acc = acc + dense[:, s]
    mean_rle = O.rle_encode(acc / n)
    total_mapped = rng.integers(30_000_000, 90_000_000, size=n).astype(np.float64)
    for Lval, tm, budget in ((76, None, 2 ** 31), (np.arange(1, n + 1) * 25.0, total_mapped, 2 ** 31),
                             (100, total_mapped, 4096), (76, 55e6, 20000)):
        got = dfb.railMatrix(["chr21"], {"chr21": mean_rle}, {"chr21": cols}, L_=Lval, cutoff=2.6,
                             maxClusterGap=3000, totalMapped=tm, engine=eng, deviceBytes=budget)["chr21"]
        want = O.rail_matrix(mean_rle, cols, 2.6, L=Lval, maxClusterGap=3000, totalMapped=tm)
        assert len(want["regions"]["start"]) > 50
        assert_regions_equal(got["regions"], want["regions"])
        assert np.array_equal(got["coverageMatrix"], want["coverageMatrix"])
    # dense sample input, and the regionMatrix cross-check of the reference's own test
    got = dfb.railMatrix("chr21", {"chr21": mean_rle}, {"chr21": dense.astype(np.int32)}, L_=76, cutoff=2.6,
                         maxClusterGap=3000, engine=eng)["chr21"]
    rm = dfb.regionMatrix({"chr21": cols}, cutoff=2.6, L_=76, maxClusterGap=3000, returnBP=False, engine=eng)["chr21"]
    assert_regions_equal(got["regions"], rm["regions"], keys=("start", "end", "cluster", "area", "value"))
    assert np.array_equal(got["coverageMatrix"], rm["coverageMatrix"])
    # nothing above the cutoff -> empty result (:190-192)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        none = dfb.railMatrix(["chr21"], {"chr21": mean_rle}, {"chr21": cols}, L_=76, cutoff=1e9, engine=eng)["chr21"]
    assert none["regions"] is None and none["coverageMatrix"] is None
    with pytest.raises(ValueError):
        dfb.railMatrix(["chr21"], {"chr21": mean_rle}, {"chr21": cols}, L_=76, engine=eng)


@pytest.mark.gpu
@pytest.mark.parametrize("N", [6400, 6401])
def test_view_dense_dev_equals_copy(dfb, eng, N):
    """dfb_cov_view_dense_dev reads a device matrix in the engine's layout in place (N = ld = 6400) and
    copies any other layout (6401): both must give the statistics of the copying entry point."""
    import torch
    from derfinder_b200 import _lib as L
    rng = np.random.default_rng(N)
    n = 7
    cov = rng.poisson(4.0, size=(n, N)).astype(np.int32)
    dev = torch.as_tensor(cov, device="cuda")
    e2 = dfb.Engine(0, stream=torch.cuda.current_stream().cuda_stream)
    try:
        mod = np.column_stack([np.ones(n), np.arange(n) % 2, rng.normal(size=n)])
        mod0 = mod[:, [0, 2]].copy()
        pl = np.array([N], dtype=np.int32)
        outs = []
        for fn in (e2.lib.dfb_cov_from_dense_dev, e2.lib.dfb_cov_view_dense_dev):
            h = C.c_void_p()
            L.check(fn(e2.h, n, N, L.DFB_I32, C.c_void_p(dev.data_ptr()), N, L.ptr(pl, C.c_int32), 1, 1, C.byref(h)))
            L.check(e2.lib.dfb_preprocess(e2.h, h, 32.0, None, 0))
            f = np.empty(N)
            L.check(e2.lib.dfb_fstats(e2.h, h, np.asfortranarray(mod).ctypes.data_as(L.c_f64p), 3,
                                      np.asfortranarray(mod0).ctypes.data_as(L.c_f64p), 2, 0.0,
                                      f.ctypes.data_as(L.c_f64p)))
            outs.append(f)
            e2.lib.dfb_cov_free(h)
        assert np.array_equal(outs[0], outs[1], equal_nan=True)  # constant rows give NaN (0 / 0)
        assert np.array_equal(dev.cpu().numpy(), cov)  # the borrowed matrix is never written or freed
    finally:
        e2.close()


@pytest.mark.gpu
def test_fstats_begin_wait_equals_blocking(dfb, eng):
    """dfb_fstats_begin / dfb_fstats_wait deliver the same vector as the blocking dfb_fstats, also when
    dfb_calculate_pvalues runs in between (the download proceeds on its own stream and host threads)
    and when the handle is freed without an explicit wait."""
    from derfinder_b200 import _lib as L
    rng = np.random.default_rng(808)
    n, N = 9, 3_000_000  # 24 MB of statistics: three 8 MiB stages
    cov = rng.poisson(3.0, size=(n, N)).astype(np.int32)
    mod = np.asfortranarray(np.column_stack([np.ones(n), np.arange(n) % 2, rng.normal(size=n)]))
    mod0 = np.asfortranarray(mod[:, [0, 2]])
    pl = np.array([N], dtype=np.int32)
    perm = random_perms(rng, 3, n).astype(np.int32)

    def handle():
        h = C.c_void_p()
        L.check(eng.lib.dfb_cov_from_dense(eng.h, n, N, L.DFB_I32, cov.ctypes.data_as(C.c_void_p), N,
                                           L.ptr(pl, C.c_int32), 1, 1, C.byref(h)))
        L.check(eng.lib.dfb_preprocess(eng.h, h, 32.0, None, 0))
        return h

    args = (mod.ctypes.data_as(L.c_f64p), 3, mod0.ctypes.data_as(L.c_f64p), 2, 0.0)
    h = handle()
    want = np.empty(N)
    L.check(eng.lib.dfb_fstats(eng.h, h, *args, want.ctypes.data_as(L.c_f64p)))
    got = np.full(N, -7.0)
    L.check(eng.lib.dfb_fstats_begin(eng.h, h, *args, got.ctypes.data_as(L.c_f64p)))
    rh, nh = C.c_void_p(), C.c_void_p()
    rc = L.check(eng.lib.dfb_calculate_pvalues(eng.h, h, None, *args[:4], 0.0, L.ptr(perm, C.c_int32), 3, -6.0, 6.0, 0,
                                               300, C.byref(rh), C.byref(nh), None))
    L.check(eng.lib.dfb_fstats_wait(eng.h))
    assert np.array_equal(got, want, equal_nan=True)
    if rc == 0:
        eng.lib.dfb_regions_free(rh)
        eng.lib.dfb_nulls_free(nh)
    got2 = np.full(N, -7.0)
    L.check(eng.lib.dfb_fstats_begin(eng.h, h, *args, got2.ctypes.data_as(L.c_f64p)))
    eng.lib.dfb_cov_free(h)  # waits for the pending download
    assert np.array_equal(got2, want, equal_nan=True)


@pytest.mark.gpu
def test_long_regions_warp_walk(dfb, eng):
    """Regions longer than a few mask words are walked by one warp each (32 bases per step, run heads by
    ballot) instead of one thread: same IRanges run-product order, so areas, means and widths must
    stay bit-identical to the oracle -- for long runs of equal values, for runs of length one, across
    segment boundaries, at odd start bits and at the very end of the vector."""
    rng = np.random.default_rng(4242)
    N = 60_000
    x = np.zeros(N)
    pos_ptr = 17
    truth_long = 0
    while pos_ptr < N - 10:
        width = int(rng.choice([1, 3, 31, 33, 64, 97, 129, 400, 2500, 7001]))
        width = min(width, N - pos_ptr)
        kind = rng.integers(0, 3)
        if kind == 0:    # piecewise constant (read-length-like runs): long Rle runs
            runs = np.repeat(rng.uniform(2.0, 9.0, size=width // 40 + 1), 40)[:width]
        elif kind == 1:  # every value different: every base its own run
            runs = rng.uniform(2.0, 9.0, size=width)
        else:            # short runs of 1..5
            runs = np.repeat(rng.uniform(2.0, 9.0, size=width), rng.integers(1, 6, size=width))[:width]
        x[pos_ptr:pos_ptr + width] = runs
        truth_long += width > 128
        pos_ptr += width + int(rng.integers(1, 70))
    x[N - 300:] = np.repeat(rng.uniform(2.0, 9.0, size=6), 50)  # a long region ending at the last base
    assert truth_long > 5
    position = (np.array([True, False, True, False, True]), np.array([20_011, 500, 19_989, 3000, 20_000]))
    for gap in (0, 600):
        got = dfb.findRegions(position, x, "chr1", cutoff=1.5, maxRegionGap=gap, maxClusterGap=3000, engine=eng)
        want = O.find_regions(position, x, 1.5, maxRegionGap=gap, maxClusterGap=3000)
        assert_regions_equal(got, want)
        assert (np.asarray(want["indexEnd"]) - np.asarray(want["indexStart"]) + 1).max() > 2000
    gb = dfb.findRegions(position, x, "chr1", cutoff=1.5, basic=True, engine=eng)
    ob = O.find_regions(position, x, 1.5, basic=True)
    for k in ("area", "width", "stat"):
        assert np.array_equal(gb[k], ob[k])
