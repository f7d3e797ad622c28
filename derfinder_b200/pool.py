"""Multi-GPU host logic: chromosome sharding and the one exchange step of the pipeline.

Chromosomes are independent up to the genome-wide p-value / FWER ranking of `mergeResults`
(R/mergeResults.R:216-235 pools the null areas of all chromosomes, :304-307 recomputes `.calcPval`
on the pool, :380-386 takes the per-permutation maximum area over the genome).  So every rank runs
the whole per-chromosome pipeline on its own GPU and only

    counts [world]            all-gather   (how many null regions each rank found)
    null areas [cap] padded   all-gather   (variable length -> fixed capacity)
    per-permutation max [B]   all-reduce(max)

cross NVLink.  Tensors stay where they are (CUDA with NCCL, CPU with gloo in the tests).  Each rank
ranks its own regions (already sorted by area) against the pooled, UNSORTED null lists with a
histogram (`dfb_regions_null_hist_dev` / `dfb_regions_hist_counts_dev`): nothing is sorted for the
pooling.
"""
from __future__ import annotations

import ctypes as C


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
    Returns (segments, global_max): `segments[r]` is rank r's area list (a view into the gathered
    buffer, in the order rank r supplied it -- padding is dropped by COUNT, so a genuine NaN area
    survives) and global_max the element-wise maximum of local_max.  `torch.cat(segments)` is the
    pooled list in rank order.  Without an initialised process group this is the identity.
    """
    import torch
    if dist is None:
        import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return [local_area], local_max
    world = dist.get_world_size(group)
    dev = local_area.device
    count = torch.tensor([local_area.numel()], device=dev, dtype=torch.int64)
    counts = torch.empty(world, device=dev, dtype=torch.int64)
    dist.all_gather_into_tensor(counts, count, group=group)
    gmax = local_max.clone()
    dist.all_reduce(gmax, op=dist.ReduceOp.MAX, group=group)
    counts_h = counts.tolist()  # the one host synchronisation of the exchange
    cap = max(max(counts_h), 1)
    mine = torch.empty(cap, device=dev, dtype=torch.float64)
    mine[: local_area.numel()] = local_area
    pool = torch.empty(world * cap, device=dev, dtype=torch.float64)
    dist.all_gather_into_tensor(pool, mine, group=group)
    return [pool[r * cap: r * cap + counts_h[r]] for r in range(world)], gmax


def genomewide_pvalues(engine, regions_handle, nulls_handle, B, cutoff_used, torch, device, dist=None,
                       merge_area=False):
    """This rank's regions ranked against the genome-wide null pool (GPU path; used by bench.py).
    Returns (pvalues, fwer) device tensors (regions in up-then-dn order) and the pool size.
    merge_area=True pools `abs(stat) * width` exactly as mergeResults does (R/mergeResults.R:231);
    False pools the summed areas calculatePvalues ranks (the two differ in the last ulps)."""
    from . import _lib as L
    lib = engine.lib
    M = int(lib.dfb_nulls_count(nulls_handle))
    R = int(lib.dfb_regions_count(regions_handle))
    area = torch.empty(max(M, 1), device=device, dtype=torch.float64)
    perm = torch.zeros(max(M, 1), device=device, dtype=torch.int32)
    mx = torch.empty(B, device=device, dtype=torch.float64)
    a_dev = torch.empty(max(R, 1), device=device, dtype=torch.float64)
    L.check(lib.dfb_nulls_copy_dev(nulls_handle, None if merge_area else C.c_void_p(area.data_ptr()),
                                   C.c_void_p(perm.data_ptr())))
    if merge_area:
        L.check(lib.dfb_nulls_merge_areas_dev(nulls_handle, C.c_void_p(area.data_ptr())))
    L.check(lib.dfb_perm_max_dev(engine.h, C.c_void_p(area.data_ptr()), C.c_void_p(perm.data_ptr()), M, B, -1.0,
                                 C.c_void_p(mx.data_ptr())))
    L.check(lib.dfb_regions_copy_areas_dev(engine.h, regions_handle, C.c_void_p(a_dev.data_ptr())))
    # the engine and torch share one stream in bench.py; otherwise order them explicitly
    shared = engine.stream_handle == torch.cuda.current_stream().cuda_stream
    if not shared:
        engine.sync()
    segments, gmax = exchange_nulls(area[:M], mx, dist=dist)
    # a permutation without any null region anywhere contributes cutoffFstatUsed (R/mergeResults.R:382)
    gmax = torch.where(gmax < 0, torch.full_like(gmax, float(cutoff_used)), gmax)
    # this rank's regions are sorted by area already: the pooled nulls (unsorted, rank by rank) are
    # ranked against them with a histogram -- nothing is sorted for the pooling
    hist = torch.zeros(R + 1, device=device, dtype=torch.int32)
    counts = torch.zeros(max(R, 1), device=device, dtype=torch.int64)
    total = 0
    if not shared:
        torch.cuda.current_stream().synchronize()
    for seg in segments:
        total += seg.numel()
        if seg.numel():
            L.check(lib.dfb_regions_null_hist_dev(engine.h, regions_handle, C.c_void_p(seg.data_ptr()), seg.numel(),
                                                  C.c_void_p(hist.data_ptr())))
    L.check(lib.dfb_regions_hist_counts_dev(engine.h, regions_handle, C.c_void_p(hist.data_ptr()),
                                            C.c_void_p(counts.data_ptr())))
    if not shared:
        engine.sync()
    a_dev, counts = a_dev[:R], counts[:R]
    # every null entry is stored twice by the reference (R/calculatePvalues.R:346-354): weight 2
    # divide by TENSORS: torch turns `tensor / python_scalar` into a multiplication by the reciprocal,
    # which is not the correctly rounded quotient R computes
    num = 2.0 * counts.double() + 1.0
    p_dev = num / torch.full_like(num, 2.0 * total + 1.0)
    fnum = ((gmax[None, :] > a_dev[:, None]).sum(dim=1) + 1).double()
    fwer = fnum / torch.full_like(fnum, float(B + 1))
    return p_dev, fwer, total
