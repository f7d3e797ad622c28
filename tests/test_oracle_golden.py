"""Pin the oracle (oracle/derfinder_oracle.py, oracle/rrng.py) to the reference's own goldens.

Mirrors the reference's testthat files: tests/testthat/test_Fstats.R, test_findRegions.R,
test_data.R (filterData known answers).  CPU only.
"""
import numpy as np
import pytest

from oracle import derfinder_oracle as O
from derfinder_b200.rrng import RRandom, permutation_indices


def test_models_rebuilt_from_coverage(golden):
    # tests/testthat/test_Fstats.R:4-40 -- sampleDepth(probs=0.5) + makeModels
    sd = O.sample_depth(O.collapse_full_coverage(golden.cov_rle), 0.5)
    mod = np.column_stack([np.ones(31), (golden.pop == "YRI") * 1.0, sd, (golden.gender == "male") * 1.0])
    assert np.array_equal(mod, golden.mod)
    assert np.array_equal(mod[:, [0, 2, 3]], golden.mod0)


def test_preprocess_matches_reference(golden):
    # tests/testthat/test_findRegions.R:21-45
    prep = O.preprocess_coverage(golden.cov, golden.position, groupInfo=golden.pop, scalefac=32, chunksize=1e3)
    assert prep["mclapplyIndex"] == [1000, 434]
    assert np.array_equal(prep["coverageProcessed"], np.log2(golden.cov + 32.0))
    assert np.array_equal(prep["meanCoverage"], golden["chr21_prep_meanCoverage"])
    assert abs(prep["meanCoverage"].sum() - 357.838709677) < 1e-8  # tests/testthat/test_data.R:38
    for k in ("CEU", "YRI"):
        assert np.array_equal(prep["groupMeans"][k], golden[f"chr21_prep_groupMeans_{k}"])


def test_chunk_lengths():
    # R/preprocessCoverage.R:182-188 incl. the exact-multiple fix
    assert O.chunk_lengths(1434, 1e3) == [1000, 434]
    assert O.chunk_lengths(2000, 1000) == [1000, 1000]
    assert O.chunk_lengths(999, 1000) == [999]
    assert O.chunk_lengths(1000, 1000) == [1000]
    assert O.chunk_lengths(10, None, mc_cores=3) == [4, 4, 2]


def test_fstats_golden(golden):
    # tests/testthat/test_Fstats.R:52-68 (expect_equivalent tolerance 1.5e-8); north_star: 1e-6 rel
    prep = O.preprocess_coverage(golden.cov, golden.position, scalefac=32, chunksize=1e3)
    F = O.calculate_stats(prep, golden.mod, golden.mod0)
    assert np.max(np.abs(F - golden.fstats) / np.abs(golden.fstats)) < 1e-7
    assert np.max(np.abs(F - golden.fstats_matrix) / np.abs(golden.fstats)) < 1e-7


def _tie_band_check(p, pg, areas, nullareas, band=1e-9):
    """H4 / SURVEY A.7: counts may differ from the R run only by nulls within `band` (relative) of
    the observed area."""
    M = len(nullareas)
    for a, x, y in zip(areas, p, pg):
        if x == y:
            continue
        near = np.sum(np.abs(nullareas - a) <= band * a)
        assert near > 0, "p-value differs without a near-tie"
        assert abs(round(x * (M + 1)) - round(y * (M + 1))) <= near


def test_calculate_pvalues_golden(golden):
    # tests/testthat/test_Fstats.R:80-87: nPermute=10, seeds=1:10, cutoff=1, RNGversion("3.5.0")
    prep = O.preprocess_coverage(golden.cov, golden.position, groupInfo=golden.pop, scalefac=32, chunksize=1e3)
    perm = permutation_indices(range(1, 11), 31, "Rounding")
    res = O.calculate_pvalues(prep, golden.mod, golden.mod0, golden.fstats, perm, cutoff=1)
    r = res["regions"]
    G = lambda k: golden["genomeRegions_" + k]
    assert len(r["start"]) == 33
    for k in ("start", "indexStart", "indexEnd", "cluster", "clusterL", "area", "value", "meanCoverage",
              "meanCEU", "meanYRI"):
        assert np.array_equal(r[k], G(k)), k
    assert np.array_equal(r["end"] - r["start"] + 1, G("width"))
    assert np.array_equal(r["log2FoldChangeYRIvsCEU"], G("log2FoldChangeYRIvsCEU"), equal_nan=True)
    assert list(r["name"]) == list(G("names"))
    assert len(res["nullStats"]) == 540
    assert np.array_equal(res["nullWidths"], G("nullWidths"))
    assert np.array_equal(res["nullPermutation"], G("nullPermutation"))
    assert np.max(np.abs(res["nullStats"] - G("nullStats")) / G("nullStats")) < 1e-10
    assert np.sum(r["pvalues"] == G("pvalues")) >= 31
    _tie_band_check(r["pvalues"], G("pvalues"), r["area"], res["nullAreas"])
    assert np.array_equal(r["significant"], G("significant"))


def test_bundled_chr21_run(golden):
    # inst/extdata/chr21: analyzeChr(..., seeds=20140330, nPermute=1, cutoffFstat=1 manual, method Matrix)
    prep = O.preprocess_coverage(golden.cov, golden.position, groupInfo=golden.pop)
    perm = permutation_indices([20140330], 31, "Rounding")
    res = O.calculate_pvalues(prep, golden.mod, golden.mod0, golden.fstats_matrix, perm, cutoff=1)
    r = res["regions"]
    for k in ("start", "indexStart", "indexEnd", "cluster", "clusterL", "area", "value", "pvalues",
              "meanCoverage"):
        assert np.array_equal(r[k], golden["chr21_regions_" + k]), k
    assert np.array_equal(res["nullWidths"], golden["chr21_regions_nullWidths"])
    assert np.max(np.abs(res["nullStats"] - golden["chr21_regions_nullStats"])) < 1e-9


def test_literal_known_answers(golden):
    # tests/testthat/test_Fstats.R:72-74, 89-90
    assert np.allclose(O.calc_pval([1, 2], [1, 2, 3, 4]), [0.8, 0.6], rtol=0, atol=0)
    assert np.array_equal(O.calculate_fwer(10, [1, 2, 3, 4], [1, 1, 3, 3], 3, [1, 2]), [1.0, 0.75])
    assert O.quantile_type7(np.repeat([1.0, 2.0], 10), 0.5) == 1.5
    assert abs(O.calc_fstat_cutoff("theoretical", 0.05, mod=golden.mod, mod0=golden.mod0) - 4.210008468) < 1e-8
    assert O.calc_fstat_cutoff("manual", 1) == 1
    # tests/testthat/test_findRegions.R:5-18
    pos = (np.array([1, 0, 1, 0, 1], bool), np.array([1000, 1000, 1000, 200, 1000]))
    ids, lens = O.cluster_maker_rle(pos)
    assert list(ids) == [1, 2] and list(lens) == [1000, 2000]
    st, en = O.cluster_maker_rle(pos, ranges=True)
    assert list(st) == [1, 1001] and list(en) == [1000, 3000]
    ids, lens = O.cluster_maker_rle(pos, maxGap=999)
    assert list(ids) == [1, 2] and list(lens) == [1000, 2000]
    ids, lens = O.cluster_maker_rle(pos, maxGap=1000)
    assert list(ids) == [1] and list(lens) == [3000]
    x = np.repeat([-1.0, 0, 1, 2], 1000)
    s = O.get_segments(x, 0.5)
    assert (list(s["upIndex"][0]), list(s["upIndex"][1])) == ([2001], [4000])
    assert (list(s["dnIndex"][0]), list(s["dnIndex"][1])) == ([1], [1000])
    s = O.get_segments(x + 3, [1, 1])
    assert (list(s["upIndex"][0]), list(s["upIndex"][1])) == ([1], [4000]) and len(s["dnIndex"][0]) == 0


def test_find_regions_known_answers(golden):
    # tests/testthat/test_findRegions.R:49-68
    F = golden.fstats
    small_pos = (np.array([False, True, False, True]), np.array([47407536, 32, 1256, 36]))
    rb = O.find_regions(small_pos, F[:68], 1, basic=True)
    assert list(rb["width"]) == [6, 15, 36]
    m = F[:68] > 1
    sums = [F[:68][m][:6].sum(), F[:68][m][6:21].sum(), F[:68][m][21:].sum()]
    assert np.allclose(rb["area"], sums, rtol=1e-14)
    assert np.allclose(rb["stat"], rb["area"] / rb["width"], rtol=1e-15)
    full = O.find_regions(small_pos, F[:68], 1)
    assert list(full["end"] - full["start"] + 1) == [6, 15, 36]
    one = O.find_regions(golden.position, F, 0, maxRegionGap=1e5, maxClusterGap=1e6)
    assert list(one["start"]) == [47407537] and list(one["end"]) == [47418068] and list(one["name"]) == ["up"]
    assert O.find_regions((np.array([False]), np.array([1])), F, 1) is None
    # default cutoff quantile(0.99): 15 candidate bases -> regions are subsets of the cutoff-1 ones
    regs = O.find_regions(golden.position, F, O.quantile_type7(F, 0.99))
    assert regs is not None and np.all(regs["value"] >= O.quantile_type7(F, 0.99))


def test_filter_data_known_answers(golden):
    # tests/testthat/test_data.R:174-191
    f = O.filter_data(golden.raw_rle, cutoff=0.5, filter="mean")
    assert f.coverage.shape == (273, 31)
    f = O.filter_data(golden.raw_rle, cutoff=10)
    assert list(f.position[0]) == [False] and list(f.position[1]) == [48129895]
    f = O.filter_data(golden.raw_rle, cutoff=0, returnMean=True)
    assert np.array_equal(f.coverage, golden.cov)
    assert np.array_equal(f.position[0], golden.position[0]) and np.array_equal(f.position[1], golden.position[1])
    assert f.coverage.shape[0] == int(np.sum(f.position[1][f.position[0]]))
    f = O.filter_data(golden.raw_rle, cutoff=-1, returnCoverage=False)
    assert list(f.position[0]) == [True] and list(f.position[1]) == [48129895]
    with pytest.raises(ValueError):
        O.filter_data(golden.raw_rle, filter="two")
    # re-filter with a previous index (R/filterData.R:159-164, example :85-88)
    f1 = O.filter_data(golden.raw_rle, cutoff=0)
    cols2 = [O.rle_encode(f1.coverage[:, s]) for s in range(2)]
    f2 = O.filter_data(cols2, cutoff=1, index=f1.position)
    dense = O.rle_decode(*f2.position)
    assert dense.sum() == f2.coverage.shape[0] and len(dense) == 48129895
    keep = (f1.coverage[:, 0] > 1) | (f1.coverage[:, 1] > 1)
    assert np.array_equal(f2.coverage, f1.coverage[keep][:, :2])


def test_r_rng_streams():
    # set.seed(1); runif(3) and sample(10) under both sample kinds (values of a stock R session)
    r = RRandom(1)
    u = [r.unif_rand() for _ in range(3)]
    assert np.allclose(u, [0.2655086631421, 0.3721238996368, 0.5728533633519], atol=1e-12)
    assert list(RRandom(1, "Rounding").sample(10) + 1) == [3, 4, 5, 7, 2, 8, 9, 6, 10, 1]
    assert list(RRandom(1, "Rejection").sample(10) + 1) == [9, 4, 7, 1, 2, 5, 3, 10, 6, 8]
    p = permutation_indices([1, 2, 3], 31, "Rejection")
    assert p.shape == (3, 31) and all(sorted(row) == list(range(31)) for row in p)


def test_rail_matrix_oracle_literal():
    """railMatrix restatement (R/railMatrix.R:176-290) on a case small enough to do by hand."""
    mean = (np.array([0.0, 6.0, 0.0, 8.0, 0.0]), np.array([3, 4, 5, 2, 6]))  # ERs 4-7 and 13-14 at cutoff 5
    s1 = (np.array([0, 3, 9, 0, 4, 0], dtype=np.int32), np.array([3, 2, 2, 5, 2, 6]))
    s2 = (np.array([1, 12, 0], dtype=np.int32), np.array([4, 9, 7]))
    res = O.rail_matrix(mean, [s1, s2], cutoff=5, L=2, maxClusterGap=300)
    assert list(res["regions"]["start"]) == [4, 13] and list(res["regions"]["end"]) == [7, 14]
    assert np.array_equal(res["coverageMatrix"], np.array([[(3 * 2 + 9 * 2) / 2, (1 + 12 * 3) / 2],
                                                           [4 * 2 / 2, 12 / 2]]))
    # normalised: x / (totalMapped / targetSize), one product per run, per-sample L
    res = O.rail_matrix(mean, [s1, s2], cutoff=5, L=[2, 4], totalMapped=[60e6, 30e6], targetSize=40e6)
    m1, m2 = 60e6 / 40e6, 30e6 / 40e6
    want = np.array([[((3 / m1) * 2 + (9 / m1) * 2) / 2, ((1 / m2) * 1 + (12 / m2) * 3) / 4],
                     [((4 / m1) * 2) / 2, ((12 / m2) * 1) / 4]])
    assert np.array_equal(res["coverageMatrix"], want)
    assert O.rail_matrix(mean, [s1, s2], cutoff=50)["regions"] is None


def test_rle_window_host_marshalling():
    """api._rle_window: the Rle cut down to the region union equals the dense gather."""
    from derfinder_b200 import api
    rng = np.random.default_rng(5)
    lens = rng.integers(1, 30, size=400)
    vals = rng.integers(0, 5, size=400)
    dense = np.repeat(vals, lens)
    cuts = np.sort(rng.choice(np.arange(1, len(dense)), size=80, replace=False))
    st, en = cuts[0::2] + 1, cuts[1::2]
    v, l = api._rle_window(vals, lens, st, en)
    want = np.concatenate([dense[a - 1:b] for a, b in zip(st, en)])
    assert np.array_equal(np.repeat(v, l), want)
    v, l = api._rle_window(vals, lens, np.array([1]), np.array([len(dense)]))
    assert np.array_equal(np.repeat(v, l), dense)


def test_merge_results_brute_force():
    """oracle.merge_results (R/mergeResults.R:199-328) against a direct restatement: p = (#{pooled null area >
    region area} + 1) / (#nulls + 1) with area = abs(stat) * width (:231), FWER from the per-permutation maxima over
    the whole genome (:380-386, an empty permutation counts as cutoffFstatUsed), both tables in decreasing-area
    order with stable ties (:236, :300)."""
    from oracle import derfinder_oracle as O
    rng = np.random.default_rng(17)
    B, cut = 6, 1.5

    def fake(R, M):
        if R == 0:
            return dict(regions=None, nullStats=None, nullWidths=None, nullPermutation=None)
        area = np.round(rng.random(R) * 12, 1)  # ties on purpose
        return dict(regions=dict(area=area, start=np.arange(R) * 10 + 1, value=area / 3.0),
                    nullStats=np.round(rng.normal(size=M) * 3, 1), nullWidths=rng.integers(1, 4, M),
                    nullPermutation=rng.integers(1, B, M))  # permutation B never occurs: the empty case
    res = {"chr1": fake(9, 14), "chr2": fake(0, 0), "chr3": fake(5, 8)}
    out = O.merge_results(res, B, cutoffFstatUsed=cut)
    fr, fn = out["fullRegions"], out["fullNullSummary"]
    z = np.concatenate([np.abs(res[c]["nullStats"]) * res[c]["nullWidths"] for c in ("chr1", "chr3")])
    zp = np.concatenate([res[c]["nullPermutation"] for c in ("chr1", "chr3")])
    a_all = np.concatenate([res[c]["regions"]["area"] for c in ("chr1", "chr3")])
    chr_all = np.array(["chr1"] * 9 + ["chr3"] * 5)
    o = sorted(range(len(a_all)), key=lambda i: (-a_all[i], i))
    assert np.array_equal(fr["area"], a_all[o]) and list(fr["seqnames"]) == list(chr_all[o])
    assert np.array_equal(fr["pvalues"], np.array([((z > a).sum() + 1) / (len(z) + 1) for a in a_all[o]]))
    mx = np.array([z[zp == b].max() if (zp == b).any() else cut for b in range(1, B + 1)])
    assert np.array_equal(fr["fwer"], np.array([((mx > a).sum() + 1) / (B + 1) for a in a_all[o]]))
    oz = sorted(range(len(z)), key=lambda i: (-z[i], i))
    assert np.array_equal(fn["area"], z[oz]) and np.array_equal(fn["permutation"], zp[oz])
    assert list(fn["chr"]) == list(np.array(["chr1"] * 14 + ["chr3"] * 8)[oz])


def test_max_region_gap_vs_region_finder(golden):
    """tests/testthat/test-maxRegionGap.R:14-79: regions of the golden F-statistics turned into a coverage track,
    filterData(cutoff = 0), then findRegions(..., cutoff = 0, maxRegionGap = 0 / 50) must equal
    bumphunter::regionFinder(x, pos, cutoff = 0, maxGap = 1 / 50) in start, end, value, area, indexStart, indexEnd.
    regionFinder is restated here independently of the oracle: clusters of positions whose neighbours are at most
    maxGap apart, inside them maximal runs of x > cutoff; value = mean, area = |sum|."""
    from oracle import derfinder_oracle as O
    prep = O.preprocess_coverage(golden.cov, golden.position, scalefac=32, chunksize=1e3)
    regs = O.find_regions(prep["position"], golden.fstats, O.quantile_type7(golden.fstats, 0.99))
    # GenomicRanges::coverage(gr): 1 on every base covered by a region (the regions are disjoint)
    L = 48129895
    starts, ends = np.asarray(regs["start"]), np.asarray(regs["end"])
    o = np.argsort(starts)
    vals, lens, cur = [], [], 0
    for s, e in zip(starts[o], ends[o]):
        if s - 1 > cur:
            vals.append(0)
            lens.append(s - 1 - cur)
        vals.append(1)
        lens.append(e - s + 1)
        cur = e
    vals.append(0)
    lens.append(L - cur)
    filt = O.filter_data([(np.asarray(vals, dtype=np.int32), np.asarray(lens, dtype=np.int64))], cutoff=0, returnMean=True)
    x = np.asarray(filt.meanCoverage, dtype=np.float64)
    pos = np.flatnonzero(np.repeat(np.asarray(filt.position[0]).astype(bool), filt.position[1])) + 1  # which(position)
    assert len(pos) == len(x)

    def region_finder(x, pos, cutoff, maxGap):
        out = []
        brk = np.flatnonzero(np.diff(pos) > maxGap) + 1  # cluster breaks
        bounds = np.concatenate(([0], brk, [len(pos)]))
        for a, b in zip(bounds[:-1], bounds[1:]):
            above = x[a:b] > cutoff
            i = a
            while i < b:
                if above[i - a]:
                    j = i
                    while j + 1 < b and above[j + 1 - a]:
                        j += 1
                    seg = x[i:j + 1]
                    out.append((pos[i], pos[j], seg.mean(), abs(seg.sum()), i + 1, j + 1))
                    i = j + 1
                else:
                    i += 1
        return sorted(out)

    for gap, bump_gap in ((0, 1), (50, 50)):
        got = O.find_regions(filt.position, x, 0, maxRegionGap=gap)
        want = region_finder(x, pos, 0, bump_gap)
        assert len(want) == len(got["start"]) > 0
        for k, col in enumerate(("start", "end", "value", "area", "indexStart", "indexEnd")):
            assert np.array_equal(np.asarray(got[col], dtype=np.float64), np.asarray([w[k] for w in want], dtype=np.float64)), (gap, col)
        assert set(got["name"]) == {"up"}
