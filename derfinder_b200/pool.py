"""Multi-GPU host logic: chromosome sharding and the one exchange step of the pipeline.

Chromosomes are independent up to the genome-wide p-value / FWER ranking of `mergeResults`
(R/mergeResults.R:216-235 pools the null areas of all chromosomes, :304-307 recomputes `.calcPval`
on the pool, :380-386 takes the per-permutation maximum area over the genome).  So every rank runs
the whole per-chromosome pipeline on its own GPU and only

    counts [world]                     all-gather on a SIDE stream (how many null regions each rank
                                       found; the host needs the largest to size the records, and
                                       waits for its peers only -- not for its own pending GPU work)
    records [world][1 + B + cap]       ONE all-gather: {count, per-permutation max area, null areas
                                       padded with -1 to the largest count}

cross NVLink.  Tensors stay where they are (CUDA with NCCL, CPU with gloo in the tests).  Each rank
ranks its own regions (already sorted by area) against the gathered, UNSORTED null areas with a
histogram (`dfb_pool_rank_dev`): nothing is sorted for the pooling and the per-permutation maxima
are reduced from the same records (no separate all-reduce).
"""
from __future__ import annotations

import ctypes as C
import os

# developer instrumentation (DFB_POOL_PROFILE=1): CUDA-event time of the sections of
# genomewide_pvalues, summed over calls -> pool.PROFILE
PROFILE = {"calls": 0, "events": []}


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


def exchange_counts(count, device, dist=None, group=None, side_stream=None):
    """Every rank's null-region count, as a Python list.  On CUDA the tiny all-gather and the host
    read run on `side_stream`, so the caller's main stream (still finishing the chromosome pass) is
    neither waited for nor drained; the host blocks until the peers have reached the same point."""
    import torch
    if dist is None:
        import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return [int(count)]
    world = dist.get_world_size(group)
    if device.type != "cuda":
        mine = torch.tensor([int(count)], dtype=torch.int64)
        out = torch.empty(world, dtype=torch.int64)
        dist.all_gather_into_tensor(out, mine, group=group)
        return out.tolist()
    side = side_stream or torch.cuda.Stream(device=device)
    with torch.cuda.stream(side):
        mine = torch.tensor([int(count)], dtype=torch.int64).pin_memory().to(device, non_blocking=True)
        out = torch.empty(world, device=device, dtype=torch.int64)
        dist.all_gather_into_tensor(out, mine, group=group)
        return out.tolist()  # synchronises the side stream only


def gather_records(record, dist=None, group=None):
    """All-gather of equal-length 1-d records -> [world, len] (the identity without a process group)."""
    import torch
    if dist is None:
        import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return record[None, :]
    world = dist.get_world_size(group)
    out = torch.empty(world * record.numel(), device=record.device, dtype=record.dtype)
    dist.all_gather_into_tensor(out, record.contiguous(), group=group)
    return out.view(world, record.numel())


_side_streams = {}


def genomewide_pvalues(engine, regions_handle, nulls_handle, B, cutoff_used, torch, device, dist=None,
                       merge_area=False):
    """This rank's regions ranked against the genome-wide null pool (GPU path; used by bench.py).
    Returns (pvalues, fwer) device tensors (regions in up-then-dn order) and the pool size.
    merge_area=True pools `abs(stat) * width` exactly as mergeResults does (R/mergeResults.R:231);
    False pools the summed areas calculatePvalues ranks (the two differ in the last ulps).

    `exchange_counts` (side stream), `dfb_pool_local_dev` (this rank's record), `gather_records`
    (one all-gather), `dfb_pool_rank_dev` (histogram of the pooled, UNSORTED nulls against this rank's
    sorted regions, p-values and FWER).  The main stream is never synchronised with the host."""
    from . import _lib as L
    lib = engine.lib
    M = int(lib.dfb_nulls_count(nulls_handle))
    R = int(lib.dfb_regions_count(regions_handle))
    marks = []

    def mark(name):
        if os.environ.get("DFB_POOL_PROFILE"):
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            marks.append((name, ev))

    mark("start")
    key = (device.index, id(dist))
    if key not in _side_streams:
        _side_streams[key] = torch.cuda.Stream(device=device)
    counts = exchange_counts(M, device, dist=dist, side_stream=_side_streams[key])
    world = len(counts)
    cap = max(max(counts), 1)
    record = torch.empty(1 + B + cap, device=device, dtype=torch.float64)
    p_dev = torch.empty(max(R, 1), device=device, dtype=torch.float64)
    fwer = torch.empty(max(R, 1), device=device, dtype=torch.float64)
    # the engine and torch share one stream in bench.py; otherwise order them explicitly
    shared = engine.stream_handle == torch.cuda.current_stream().cuda_stream
    if not shared:
        torch.cuda.current_stream().synchronize()
    L.check(lib.dfb_pool_local_dev(engine.h, nulls_handle, B, 1 if merge_area else 0, cap,
                                   C.c_void_p(record.data_ptr())))
    if not shared:
        engine.sync()
    mark("local")
    records = gather_records(record, dist=dist)
    mark("exchange")
    if not shared:
        torch.cuda.current_stream().synchronize()
    L.check(lib.dfb_pool_rank_dev(engine.h, regions_handle, C.c_void_p(records.data_ptr()), world, B, cap,
                                  float(cutoff_used), C.c_void_p(p_dev.data_ptr()), C.c_void_p(fwer.data_ptr())))
    if not shared:
        engine.sync()
    mark("rank")
    if marks:
        PROFILE["calls"] += 1
        PROFILE["events"].append(marks)
    return p_dev[:R], fwer[:R], int(sum(counts))


def reset_profile():
    PROFILE["calls"] = 0
    PROFILE["events"].clear()


def profile_summary():
    """Mean milliseconds per section over the recorded calls (call after a device synchronise)."""
    out = {}
    for marks in PROFILE["events"]:
        for (_, a), (name, b) in zip(marks[:-1], marks[1:]):
            out[name] = out.get(name, 0.0) + a.elapsed_time(b)
    return {k: v / max(PROFILE["calls"], 1) for k, v in out.items()}
