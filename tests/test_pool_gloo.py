"""world_size-2/3 gloo tests (CPU) of the multi-GPU host logic: chromosome sharding and the exchange
protocol of the genome-wide p-value / FWER ranking (R/mergeResults.R:216-235, 304-307, 380-386):
counts, region records, histogram all-reduce.  The result is checked against the oracle on the
concatenation of every rank's nulls."""
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
        M = [13, 0, 5][rank % 3] if rank else 13  # ragged, including a rank without nulls for world 3
        R = [3, 5, 0][rank % 3]                   # ... and one without regions
        if world == 2:
            M, R = [13, 4][rank], [3, 5][rank]
        area = np.round(rng.gamma(2.0, 3.0, size=M), 1)  # rounded: ties between nulls and regions happen
        perm = rng.integers(1, B - 1, size=M)  # permutations B-1, B never produce a region anywhere
        regions = np.sort(np.round(rng.gamma(2.0, 3.0, size=R), 1))[::-1]
        mx = np.full(B, -1.0)
        for a, p in zip(area, perm):
            mx[p - 1] = max(mx[p - 1], a)
        # the protocol of pool.genomewide_pvalues with numpy standing in for the three library calls
        counts = pool.exchange_counts([R], torch.device("cpu"))  # first call only: agree on a capacity
        assert counts[rank] == [R] and len(counts) == world
        rcap = max(c[0] for c in counts) + 2
        record = np.concatenate([[R, M], mx, regions, np.full(rcap - R, -1.0)])   # dfb_pool_local_dev
        records = pool.gather_records(torch.as_tensor(record)).numpy()
        assert records.shape == (world, 2 + B + rcap)
        total = records[:, 1].sum()
        lists = records[:, 2 + B:]                                                 # dfb_pool_hist_dev: merge
        n = world * rcap
        merged, src = np.empty(n), np.empty(n, dtype=np.int64)
        for w in range(world):
            for i in range(rcap):
                a = lists[w, i]
                pos = i + sum(int(np.sum(lists[w2] >= a)) for w2 in range(w)) \
                    + sum(int(np.sum(lists[w2] > a)) for w2 in range(w + 1, world))
                merged[pos], src[pos] = a, w * rcap + i
        assert np.all(np.diff(merged) <= 0) and sorted(src) == list(range(n))  # a stable merge, no sort
        hist = np.zeros(world * rcap + 1, dtype=np.int64)
        for v in area:
            hist[int(np.sum(merged >= v))] += 1
        hist = pool.reduce_histogram(torch.as_tensor(hist)).numpy()
        incl = np.cumsum(hist[:-1])                                                # dfb_pool_rank_dev
        gmax = records[:, 2:2 + B].max(axis=0)
        gmax = np.where(gmax < 0, 3.5, gmax)  # empty permutation -> cutoffFstatUsed
        pv, fw = np.zeros(R), np.zeros(R)
        for i, sidx in enumerate(src):
            w, il = divmod(int(sidx), rcap)
            if w == rank and il < R:
                pv[il] = (2.0 * incl[i] + 1.0) / (2.0 * total + 1.0)
                fw[il] = (np.sum(gmax > merged[i]) + 1) / (B + 1)
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), area=area, perm=perm, regions=regions, pv=pv, fw=fw,
                 records=records)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_pooled_ranking_protocol_gloo(tmp_path, world):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    res = [np.load(tmp_path / f"r{r}.npz") for r in range(world)]
    B = 7
    for r in res:
        np.testing.assert_array_equal(r["records"], res[0]["records"])  # the same buffer on every rank
    # every rank's pooled p-values / FWER == the reference's mergeResults arithmetic on the concatenation of
    # all ranks' nulls (.calcPval on the duplicated pool, .calculateFWER)
    all_area = np.concatenate([r["area"] for r in res])
    all_perm = np.concatenate([r["perm"] for r in res])
    for r in res:
        if len(r["regions"]) == 0:
            continue
        np.testing.assert_array_equal(r["pv"], O.calc_pval(r["regions"], np.concatenate([all_area, all_area])))
        np.testing.assert_array_equal(r["fw"], O.calculate_fwer(3.5, all_area, all_perm, B, r["regions"]))


def test_exchange_is_identity_without_process_group():
    import torch
    rec = torch.arange(4, dtype=torch.float64)
    assert pool.exchange_counts([3, 9], torch.device("cpu")) == [[3, 9]]
    out = pool.gather_records(rec)
    assert out.shape == (1, 4) and out.data_ptr() == rec.data_ptr()
    h = torch.arange(3)
    assert pool.reduce_histogram(h) is h
