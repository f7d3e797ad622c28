"""Multi-GPU host logic: chromosome sharding and the one exchange step of the pipeline.

Chromosomes are independent up to the genome-wide p-value / FWER ranking of `mergeResults`
(R/mergeResults.R:216-235 pools the null areas of all chromosomes, :304-307 recomputes `.calcPval`
on the pool, :380-386 takes the per-permutation maximum area over the genome).  So every rank runs
the whole per-chromosome pipeline on its own GPU and only

    counts [world]            all-gather   (how many null regions each rank found)
    null areas [cap] padded   all-gather   (variable length -> fixed capacity, NaN padded)
    per-permutation max [B]   all-reduce(max)

cross NVLink.  Tensors stay where they are (CUDA with NCCL, CPU with gloo in the tests); the
ranking itself is the library's `dfb_calc_pval_dev` / `fwer` on the GPU.
"""
from __future__ import annotations

import ctypes as C

import numpy as np


def assign_chromosomes(sizes, world):
    """Greedy longest-first bin packing of chromosomes (name -> kept bases N) onto `world` ranks.
    Returns a list of name lists, deterministic for equal sizes (ties by name)."""
    bins = [[] for _ in range(world)]
    load = [0] * world
    for name, sz in sorted(sizes.items(), key=lambda kv: (-kv[1], str(kv[0]))):
        r = min(range(world), key=lambda i: (load[i], i))
        bins[r].append(name)
        load[r] += int(sz)
    return bins


def exchange_nulls(local_area, local_max, dist=None, group=None):
    """Pool null areas and per-permutation maxima over all ranks.

    local_area: 1-d float64 tensor, this rank's null areas (any length, may be empty).
    local_max : [B] float64 tensor, max null area per permutation on this rank, NEGATIVE where the
                rank found no region for that permutation (areas are >= 0).
    Returns (pool, global_max): the concatenation of all ranks' areas in rank order, and the
    element-wise maximum of local_max.  Without an initialised process group this is the identity.
    """
    import torch
    if dist is None:
        import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return local_area, local_max
    world = dist.get_world_size(group)
    dev = local_area.device
    count = torch.tensor([local_area.numel()], device=dev, dtype=torch.int64)
    counts = torch.empty(world, device=dev, dtype=torch.int64)
    dist.all_gather_into_tensor(counts, count, group=group)
    cap = max(int(counts.max().item()), 1)
    mine = torch.full((cap,), float("nan"), device=dev, dtype=torch.float64)
    mine[: local_area.numel()] = local_area
    pool = torch.empty(world * cap, device=dev, dtype=torch.float64)
    dist.all_gather_into_tensor(pool, mine, group=group)
    # drop the padding by COUNT (a genuine NaN area, e.g. from NaN F-statistics, must survive)
    keep = (torch.arange(cap, device=dev)[None, :] < counts[:, None]).reshape(-1)
    pool = pool[keep].contiguous()
    gmax = local_max.clone()
    dist.all_reduce(gmax, op=dist.ReduceOp.MAX, group=group)
    return pool, gmax


def genomewide_pvalues(engine, regions_handle, nulls_handle, B, cutoff_used, torch, device, dist=None):
    """This rank's regions ranked against the genome-wide null pool (GPU path; used by bench.py).
    Returns (pvalues, fwer) device tensors and the pool size."""
    from . import _lib as L
    lib = engine.lib
    M = int(lib.dfb_nulls_count(nulls_handle))
    R = int(lib.dfb_regions_count(regions_handle))
    area = torch.empty(max(M, 1), device=device, dtype=torch.float64)
    perm = torch.zeros(max(M, 1), device=device, dtype=torch.int32)
    L.check(lib.dfb_nulls_copy_dev(nulls_handle, C.c_void_p(area.data_ptr()), C.c_void_p(perm.data_ptr())))
    mx = torch.empty(B, device=device, dtype=torch.float64)
    L.check(lib.dfb_perm_max_dev(engine.h, C.c_void_p(area.data_ptr()), C.c_void_p(perm.data_ptr()), M, B, -1.0,
                                 C.c_void_p(mx.data_ptr())))
    engine.sync()  # the library works on its own stream handle; torch collectives follow
    pool, gmax = exchange_nulls(area[:M], mx, dist=dist)
    # a permutation without any null region anywhere contributes cutoffFstatUsed (R/mergeResults.R:382)
    gmax = torch.where(gmax < 0, torch.full_like(gmax, float(cutoff_used)), gmax)
    a_host = np.empty(R)
    L.check(lib.dfb_regions_get(regions_handle, None, None, None, a_host.ctypes.data_as(L.c_f64p), None, None, None,
                                None))
    a_dev = torch.as_tensor(a_host, device=device)
    p_dev = torch.empty(R, device=device, dtype=torch.float64)
    torch.cuda.current_stream().synchronize()
    L.check(lib.dfb_calc_pval_dev(engine.h, C.c_void_p(a_dev.data_ptr()), R, C.c_void_p(pool.data_ptr()),
                                  pool.numel(), 2, C.c_void_p(p_dev.data_ptr())))
    fwer = ((gmax[None, :] > a_dev[:, None]).sum(dim=1) + 1).double() / (B + 1)
    return p_dev, fwer, int(pool.numel())
