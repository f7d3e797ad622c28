"""Developer probe: host wall time of every C-ABI call of the device-resident C2 step in steady state
(python profiles/step_probe.py on a GPU box).  A call that blocks shows the device time it waits for;
what remains between the final synchronisation of one step and the first kernel of the next is host
work during which the GPU has nothing queued."""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
import derfinder_b200 as dfb  # noqa: E402
from derfinder_b200 import _lib as L  # noqa: E402
from derfinder_b200.rrng import permutation_indices  # noqa: E402

cfg = bench.WORKLOADS["c2"]
dev = torch.device("cuda", 0)
work = bench.make_workload(cfg, 20140925, torch=torch, device=dev)
perm = permutation_indices(range(1, cfg["B"] + 1), cfg["n"], "Rejection")
cut = bench.theoretical_cutoff(cfg["alpha"], work["mod"], work["mod0"])
eng = dfb.Engine(0, stream=torch.cuda.current_stream().cuda_stream)
s = bench.Step(eng, work, perm, cut, torch)
lib = eng.lib
acc = {}
steps = 30
for it in range(steps + 5):
    t = [time.perf_counter()]
    names = []

    def lap(name):
        t.append(time.perf_counter())
        names.append(name)

    cov = C.c_void_p()
    L.check(lib.dfb_cov_view_dense_dev(eng.h, s.n, s.N, L.DFB_I32, C.c_void_p(work["cov"].data_ptr()), s.N,
                                       L.ptr(s.pl, C.c_int32), len(s.pl), work["pos_first"], C.byref(cov)))
    lap("cov_view_dense_dev")
    L.check(lib.dfb_preprocess(eng.h, cov, 32.0, None, 0))
    lap("preprocess")
    L.check(lib.dfb_fstats(eng.h, cov, s._fp(s.mod), s.mod.shape[1], s._fp(s.mod0), s.mod0.shape[1], 0.0, None))
    lap("fstats")
    rh, nh = C.c_void_p(), C.c_void_p()
    L.check(lib.dfb_calculate_pvalues(eng.h, cov, None, s._fp(s.mod), s.mod.shape[1], s._fp(s.mod0), s.mod0.shape[1],
                                      0.0, L.ptr(s.perm, C.c_int32), s.perm.shape[0], -s.cut, s.cut, 0, 3000,
                                      C.byref(rh), C.byref(nh), None))
    lap("calculate_pvalues")
    lib.dfb_regions_free(rh)
    lap("regions_free")
    lib.dfb_nulls_free(nh)
    lap("nulls_free")
    lib.dfb_cov_free(cov)
    lap("cov_free")
    if it >= 5:
        for n_, a, b in zip(names, t, t[1:]):
            acc[n_] = acc.get(n_, 0.0) + (b - a) * 1e3
torch.cuda.synchronize()
print({k: round(v / steps, 4) for k, v in acc.items()}, "sum", round(sum(acc.values()) / steps, 3))
