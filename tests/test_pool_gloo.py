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
        perm = rng.integers(1, B - 1, size=M)  # permutations B-1, B never produce a region anywhere
        mx = np.full(B, -1.0)
        for a, p in zip(area, perm):
            mx[p - 1] = max(mx[p - 1], a)
        # the exchange of pool.genomewide_pvalues: counts first, then ONE all-gather of the records
        # [count | per-permutation max or -1 | areas padded with -1] that dfb_pool_local_dev writes
        counts = pool.exchange_counts(M, torch.device("cpu"))
        assert counts[rank] == M and len(counts) == world
        cap = max(max(counts), 1)
        record = np.concatenate([[float(M)], mx, area, np.full(cap - M, -1.0)])
        records = pool.gather_records(torch.as_tensor(record))
        assert tuple(records.shape) == (world, 1 + B + cap)
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), area=area, perm=perm, mx=mx, records=records.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_record_exchange_gloo(tmp_path, world):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    res = [np.load(tmp_path / f"r{r}.npz") for r in range(world)]
    B = 7
    cap = max(max(len(r["area"]) for r in res), 1)
    want = np.stack([np.concatenate([[float(len(r["area"]))], r["mx"], r["area"], np.full(cap - len(r["area"]), -1.0)])
                     for r in res])
    for r in res:
        np.testing.assert_array_equal(r["records"], want)  # the same buffer on every rank
    # what dfb_pool_rank_dev computes from the records == the reference's mergeResults arithmetic on the
    # concatenation of all ranks' nulls (.calcPval on the duplicated pool, .calculateFWER)
    rec = res[0]["records"]
    all_area = np.concatenate([r["area"] for r in res])
    all_perm = np.concatenate([r["perm"] for r in res])
    gmax = rec[:, 1:1 + B].max(axis=0)
    gmax = np.where(gmax < 0, 3.5, gmax)  # empty permutation -> cutoffFstatUsed
    region_area = np.array([0.5, 4.0, 50.0])
    fwer = ((gmax[None, :] > region_area[:, None]).sum(axis=1) + 1) / (B + 1)
    np.testing.assert_array_equal(fwer, O.calculate_fwer(3.5, all_area, all_perm, B, region_area))
    pooled = rec[:, 1 + B:].ravel()
    pooled = pooled[pooled >= 0]  # pads never outrank a region
    total = rec[:, 0].sum()
    assert total == len(all_area) == len(pooled)
    cnt = (pooled[None, :] > region_area[:, None]).sum(axis=1)
    p = (2.0 * cnt + 1.0) / (2.0 * total + 1.0)
    np.testing.assert_array_equal(p, O.calc_pval(region_area, np.concatenate([all_area, all_area])))


def test_exchange_is_identity_without_process_group():
    import torch
    rec = torch.arange(4, dtype=torch.float64)
    assert pool.exchange_counts(3, torch.device("cpu")) == [3]
    out = pool.gather_records(rec)
    assert out.shape == (1, 4) and out.data_ptr() == rec.data_ptr()
