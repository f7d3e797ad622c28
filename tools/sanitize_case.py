"""Small end-to-end case for compute-sanitizer (memcheck / racecheck / synccheck): every kernel family of the
hot path on inputs small enough for the tool's 10-100x slowdown -- filterData from Rle, preprocessCoverage,
calculateStats, calculatePvalues through the tcgen05 null scan (n = 12: four accumulator stages; n = 60:
three stages, two MMA issuers with the sequence hand-over, several passes over the permutation groups),
the long-region warp walk, regionMatrix, the pooled p-values."""
import sys

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import derfinder_b200 as dfb  # noqa: E402
from derfinder_b200 import pool  # noqa: E402
from oracle import derfinder_oracle as O  # noqa: E402
from test_gpu_parity import random_perms, random_rle_cols, synth  # noqa: E402

eng = dfb.Engine(0)
rng = np.random.default_rng(5)
# K1: filter from Rle columns, "one" and "mean", with normalisation
cols = random_rle_cols(rng, 5, 60011, 37, 0.5)
for flt, norm in (("one", False), ("mean", True)):
    tm = rng.uniform(30e6, 120e6, size=5) if norm else None
    f = dfb.filterData(cols, cutoff=3.0, filter=flt, totalMapped=tm, targetSize=80e6, returnMean=True, engine=eng)
    o = O.filter_data(cols, cutoff=3.0, filter=flt, totalMapped=tm, targetSize=80e6, returnMean=True)
    assert np.array_equal(f["coverage"].to_numpy(), o.coverage)
# K2..K5: two model shapes
for n, extra, N, B, alpha in ((12, 0, 6000, 70, 0.05), (60, 3, 2600, 300, 1e-3)):
    cov, position, group, models = synth(rng, N, n, p_extra=extra, seg=[N // 2, 33, N - N // 2 - 33])
    perm = random_perms(rng, B, n)
    cut = dfb.calcFstatCutoff("theoretical", alpha, models=models)
    prep = dfb.preprocessCoverage(dict(coverage=cov, position=position), groupInfo=group, engine=eng)
    F = dfb.calculateStats(prep, models, engine=eng)
    res = dfb.calculatePvalues(prep, models, F, permIndex=perm, cutoff=cut, engine=eng)
    print("case n=%d: regions %s nulls %s" % (n, None if res["regions"] is None else len(res["regions"]["start"]),
                                              None if res["nullStats"] is None else len(res["nullStats"])), flush=True)
# K3 long regions + K6
x = np.zeros(20000)
x[100:9000] = np.repeat(rng.uniform(2, 9, size=89), 100)
x[12000:12040] = 5.0
pos = (np.array([True]), np.array([20000]))
r = dfb.findRegions(pos, x, "chr1", cutoff=1.5, maxRegionGap=0, maxClusterGap=3000, engine=eng)
assert r is not None and len(r["start"]) == 2
icols = [(v.astype(np.int32), l) for v, l in random_rle_cols(rng, 4, 30000, 40, 0.4)]
rm = dfb.regionMatrix({"chr21": icols}, cutoff=2, L=36, engine=eng)["chr21"]
assert rm["coverageMatrix"] is not None
eng.close()
print("sanitize case ok", flush=True)
