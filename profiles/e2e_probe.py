"""Developer probe: per-call wall time of the result downloads of one C2 step (python profiles/e2e_probe.py on a GPU box)."""
import sys, time, ctypes as C, numpy as np
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import torch, bench
import derfinder_b200 as dfb
from derfinder_b200 import _lib as L
from derfinder_b200.rrng import permutation_indices
cfg = bench.WORKLOADS["c2"]
dev = torch.device("cuda", 0)
work = bench.make_workload(cfg, 20140925, torch=torch, device=dev)
perm = permutation_indices(range(1, cfg["B"] + 1), cfg["n"], "Rejection")
cut = bench.theoretical_cutoff(cfg["alpha"], work["mod"], work["mod0"])
eng = dfb.Engine(0, stream=torch.cuda.current_stream().cuda_stream)
step = bench.Step(eng, work, perm, cut, torch)
lib = eng.lib
for it in range(4):
    (R, M), rh, nh, cov = step.device_pass(keep=True)
    torch.cuda.synchronize()
    st, en, ist, ien, cl = (np.empty(R, dtype=np.int32) for _ in range(5))
    val, area, cll, pv, mc = (np.empty(R) for _ in range(5))
    if it == 0:
        bufs = (np.empty(2 * M + 64), np.empty(2 * M + 64), np.empty(2 * M + 64, dtype=np.int32), np.empty(2 * M + 64, dtype=np.int32))
        for b in bufs: b[:] = 0
    t = [time.perf_counter()]
    L.check(lib.dfb_regions_get(rh, L.ptr(st, C.c_int32), L.ptr(en, C.c_int32), step._fp(val), step._fp(area), L.ptr(ist, C.c_int32), L.ptr(ien, C.c_int32), L.ptr(cl, C.c_int32), step._fp(cll))); t.append(time.perf_counter())
    L.check(lib.dfb_regions_get_pvalues(rh, step._fp(pv))); t.append(time.perf_counter())
    L.check(lib.dfb_regions_view_means(eng.h, rh, cov, -1, step._fp(mc))); t.append(time.perf_counter())
    ns, na, nw, npid = (a[:2 * M] for a in bufs)
    L.check(lib.dfb_nulls_get(nh, 1, step._fp(ns), None, None, None)); t.append(time.perf_counter())
    L.check(lib.dfb_nulls_get(nh, 1, None, None, step._fp(na), None)); t.append(time.perf_counter())
    L.check(lib.dfb_nulls_get(nh, 1, None, L.ptr(nw, C.c_int32), None, None)); t.append(time.perf_counter())
    L.check(lib.dfb_nulls_get(nh, 1, None, None, None, L.ptr(npid, C.c_int32))); t.append(time.perf_counter())
    t0 = time.perf_counter(); L.check(lib.dfb_nulls_get(nh, 0, step._fp(ns), None, None, None)); t1 = time.perf_counter()
    print("dup0 stat(9MB)", round((t1 - t0) * 1e3, 3))
    F = np.empty(step.N)
    t0 = time.perf_counter(); L.check(lib.dfb_cov_get_mean(cov, step._fp(F))); t1 = time.perf_counter()
    print("R", R, "M", M, "ms:", [round((b - a) * 1e3, 3) for a, b in zip(t, t[1:])], "get_mean(38MB)", round((t1 - t0) * 1e3, 3))
    lib.dfb_regions_free(rh); lib.dfb_nulls_free(nh); lib.dfb_cov_free(cov)
