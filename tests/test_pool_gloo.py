"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: chromosome sharding and the null-area
exchange that precedes the genome-wide p-value / FWER ranking (R/mergeResults.R:216-235, 304-307,
380-386).  The ranking itself is checked against the oracle on the pooled data."""
import os
import socket

import numpy as np
import pytest

from derfinder_b200 import pool
from oracle import derfinder_oracle as O


def test_assign_chromosomes_balanced_and_deterministic():
    sizes = {f"chr{i}": s for i, s in enumerate([249, 243, 198, 191, 181, 171, 159, 146, 141, 135, 135, 133, 115,
                                                 107, 102, 90, 81, 78, 59, 63, 48, 51, 155, 59], start=1)}
    for world in (1, 2, 4, 8):
        bins = pool.assign_chromosomes(sizes, world)
        assert sorted(sum(bins, [])) == sorted(sizes)
        loads = [sum(sizes[c] for c in b) for b in bins]
        assert max(loads) - min(loads) <= max(sizes.values())
        assert bins == pool.assign_chromosomes(dict(reversed(list(sizes.items()))), world)
    assert pool.assign_chromosomes({"a": 5}, 4) == [["a"], [], [], []]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        B = 7
        rng = np.random.default_rng(100 + rank)
        M = [13, 0, 5][rank % 3] if rank else 13  # ragged, including an empty rank for world 3
        if world == 2:
            M = [13, 4][rank]
        area = rng.gamma(2.0, 3.0, size=M)
        if rank == 0:
            area[3] = np.nan  # a NaN area must survive the padding removal
        perm = rng.integers(1, B - 1, size=M)  # permutations B-1, B never produce a region anywhere
        mx = np.full(B, -1.0)
        for a, p in zip(area, perm):
            if not np.isnan(a):
                mx[p - 1] = max(mx[p - 1], a)
        segs, gmax = pool.exchange_nulls(torch.as_tensor(area), torch.as_tensor(mx))
        assert len(segs) == world and [int(s.numel()) for s in segs][rank] == M
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), area=area, perm=perm, pool=torch.cat(segs).numpy(),
                 gmax=gmax.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_exchange_nulls_gloo(tmp_path, world):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    res = [np.load(tmp_path / f"r{r}.npz") for r in range(world)]
    want_pool = np.concatenate([r["area"] for r in res])
    B = 7
    for r in res:
        np.testing.assert_array_equal(r["pool"], want_pool)  # rank order, NaN kept, padding gone
        np.testing.assert_array_equal(r["gmax"], res[0]["gmax"])
    # genome-wide ranking on the pool == the reference's mergeResults arithmetic on the concatenation
    all_area = want_pool[~np.isnan(want_pool)]
    all_perm = np.concatenate([r["perm"] for r in res])[~np.isnan(want_pool)]
    gmax = np.where(res[0]["gmax"] < 0, 3.5, res[0]["gmax"])  # empty permutation -> cutoffFstatUsed
    region_area = np.array([0.5, 4.0, 50.0])
    fwer = ((gmax[None, :] > region_area[:, None]).sum(axis=1) + 1) / (B + 1)
    np.testing.assert_array_equal(fwer, O.calculate_fwer(3.5, all_area, all_perm, B, region_area))
    # pooled p-values from per-rank SORTED lists (no second sort) == .calcPval on the duplicated pool
    cnt = np.zeros(len(region_area), dtype=np.int64)
    for r in res:
        srt = np.sort(r["area"][~np.isnan(r["area"])])
        cnt += len(srt) - np.searchsorted(srt, region_area, side="right")
    p = (2.0 * cnt + 1.0) / (2.0 * len(all_area) + 1.0)
    np.testing.assert_array_equal(p, O.calc_pval(region_area, np.concatenate([all_area, all_area])))


def test_exchange_is_identity_without_process_group():
    import torch
    a, m = torch.arange(4, dtype=torch.float64), torch.tensor([1.0, -1.0])
    p, g = pool.exchange_nulls(a, m)
    assert len(p) == 1 and p[0] is a and g is m
