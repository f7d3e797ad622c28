"""torchrun worker of tests/test_multi_gpu.py: 3 * world - 1 synthetic chromosomes (ragged, one without
regions) bin-packed over the ranks; every rank pools its chromosomes with dfb_merge_results; rank 0 also
runs ALL chromosomes on its own GPU without a communicator and the oracle's .calcPval / .calculateFWER on
the concatenated nulls; all three must agree bit for bit."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    import derfinder_b200 as dfb
    from derfinder_b200 import pool
    from oracle import derfinder_oracle as O
    from test_gpu_parity import random_perms, synth
    from test_gpu_parity_large import _chromosome_result, _fetch

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = dfb.Engine(local)
    w, r = pool.init_communicator(eng, dist)
    assert (w, r) == (world, rank)

    n, B = 14, 24
    nchr = 3 * world - 1
    master = np.random.default_rng(4321)
    _, _, _, models = synth(master, 2000, n, p_extra=1)
    perm = random_perms(master, B, n)
    cut = dfb.calcFstatCutoff("theoretical", 0.05, models=models)
    sizes = {c: int(master.integers(2000, 9000)) for c in range(nchr)}
    seeds = {c: int(master.integers(1, 1 << 30)) for c in range(nchr)}
    empty = 1  # this chromosome gets a cutoff nothing passes: regions = NULL
    owner = {}
    for rk, names in enumerate(pool.assign_chromosomes(sizes, world)):
        for c in names:
            owner[c] = rk

    def run(engine, c):
        return _chromosome_result(dfb, engine, np.random.default_rng(seeds[c]), sizes[c], n, models, perm,
                                  1e12 if c == empty else cut)

    for merge_area in (True, False):
        mine = sorted(c for c in range(nchr) if owner[c] == rank)
        res = {c: run(eng, c) for c in mine}
        outs, total = pool.merge_results(eng, [(res[c][1], res[c][2]) for c in mine], B, cut, merge_area=merge_area)
        got = {c: outs[i] for i, c in enumerate(mine)}
        # gather everything on rank 0
        allgot = [None] * world
        dist.all_gather_object(allgot, {c: (got[c][0].tolist(), got[c][1].tolist()) for c in mine})
        if rank == 0:
            solo = dfb.Engine(local)  # no communicator: pools locally
            sres = {c: run(solo, c) for c in range(nchr)}
            souts, stotal = pool.merge_results(solo, [(sres[c][1], sres[c][2]) for c in range(nchr)], B, cut,
                                               merge_area=merge_area)
            assert stotal == total, (stotal, total)
            info = {c: (None if sres[c][1] is None else _fetch(solo, sres[c][1], sres[c][2])) for c in range(nchr)}
            key = (lambda i: np.abs(i["nstat"]) * i["nwidth"]) if merge_area else (lambda i: i["narea"])
            pool_area = np.concatenate([key(i) for i in info.values() if i is not None])
            pool_perm = np.concatenate([i["nperm"] for i in info.values() if i is not None])
            merged = {}
            for d in allgot:
                merged.update(d)
            assert sorted(merged) == list(range(nchr))
            for c in range(nchr):
                p_multi, f_multi = np.asarray(merged[c][0]), np.asarray(merged[c][1])
                assert np.array_equal(p_multi, souts[c][0]), ("p vs single GPU", c)
                assert np.array_equal(f_multi, souts[c][1]), ("fwer vs single GPU", c)
                if info[c] is None:
                    assert len(p_multi) == 0
                    continue
                assert np.array_equal(p_multi, O.calc_pval(info[c]["area"], np.repeat(pool_area, 2))), ("p vs oracle", c)
                assert np.array_equal(f_multi, O.calculate_fwer(cut, pool_area, pool_perm, B, info[c]["area"])), c
            for h, rh, nh in sres.values():
                if rh is not None:
                    solo.lib.dfb_regions_free(rh)
                    solo.lib.dfb_nulls_free(nh)
                solo.lib.dfb_cov_free(h)
            solo.close()
        for h, rh, nh in res.values():
            if rh is not None:
                eng.lib.dfb_regions_free(rh)
                eng.lib.dfb_nulls_free(nh)
            eng.lib.dfb_cov_free(h)
        dist.barrier()
    eng.close()
    if rank == 0:
        print("MERGE-PARITY-OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
