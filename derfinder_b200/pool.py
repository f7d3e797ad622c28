"""Multi-GPU host logic: chromosome sharding and the one exchange step of the pipeline.

Chromosomes are independent up to the genome-wide p-value / FWER ranking of `mergeResults`
(R/mergeResults.R:216-235 pools the null areas of all chromosomes, :304-307 recomputes `.calcPval`
on the pool, :380-386 takes the per-permutation maximum area over the genome).  So every rank runs
the whole per-chromosome pipeline on its own GPU and only

    records [world][2 + B + rcap] all-gather: {regions R, null regions M, per-permutation max null
                                  area, this rank's REGION areas (sorted, padded with -1 to rcap)}
    histogram [world * rcap + 1]  all-reduce(sum): this rank's nulls binned against the merged region
                                  list of all ranks

cross NVLink -- a few MB per rank.  The null lists themselves (the bulk: ~10 MB per chromosome at
C2, far more at C3/C4) never leave their GPU, and per-rank work is O(own nulls * log(all regions))
whatever the number of GPUs: every rank merges the gathered region areas into one descending list,
histograms its own UNSORTED nulls against it, and after the all-reduce reads the pooled count of
its own regions off the prefix sums.  The record capacity `rcap` is agreed once (a count exchange
on the first call) and kept with head room; the counts travel inside the records, so a rank that
outgrows the capacity is seen by everybody and the exchange is repeated once with a larger one.
Tensors stay where they are (CUDA with NCCL, CPU with gloo in the tests); nothing is read back to
the host until all the work of the step has been queued.
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


def exchange_counts(values, device, dist=None, group=None, side_stream=None):
    """All-gather of a few integers per rank (here: region and null-region counts) -> list of
    per-rank lists.  On CUDA the tiny collective and the host read run on `side_stream`, so the
    caller's main stream (still finishing the chromosome pass) is neither waited for nor drained;
    the host blocks until the peers have reached the same point."""
    import torch
    if dist is None:
        import torch.distributed as dist
    values = [int(v) for v in values]
    if not (dist.is_available() and dist.is_initialized()):
        return [values]
    world = dist.get_world_size(group)
    k = len(values)
    if device.type != "cuda":
        mine = torch.tensor(values, dtype=torch.int64)
        out = torch.empty(world * k, dtype=torch.int64)
        dist.all_gather_into_tensor(out, mine, group=group)
        return out.view(world, k).tolist()
    side = side_stream or torch.cuda.Stream(device=device)
    with torch.cuda.stream(side):
        mine = torch.tensor(values, dtype=torch.int64).pin_memory().to(device, non_blocking=True)
        out = torch.empty(world * k, device=device, dtype=torch.int64)
        dist.all_gather_into_tensor(out, mine, group=group)
        return out.view(world, k).tolist()  # synchronises the side stream only


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


def reduce_histogram(hist, dist=None, group=None):
    """In-place all-reduce(sum) of the int64 histogram (no-op without a process group)."""
    if dist is None:
        import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=group)
    return hist


_side_streams = {}
_capacity = {}  # (device, process group) -> agreed record capacity


def _headroom(r):
    return int(r) + (int(r) >> 2) + 1024


def genomewide_pvalues(engine, regions_handle, nulls_handle, B, cutoff_used, torch, device, dist=None,
                       merge_area=False):
    """This rank's regions ranked against the genome-wide null pool (GPU path; used by bench.py).
    Returns (pvalues, fwer) device tensors (regions in up-then-dn order) and the pool size.
    merge_area=True pools `abs(stat) * width` exactly as mergeResults does (R/mergeResults.R:231);
    False pools the summed areas calculatePvalues ranks (the two differ in the last ulps).

    dfb_pool_local_dev -> gather_records -> dfb_pool_hist_dev -> reduce_histogram ->
    dfb_pool_rank_dev, then ONE host read of the gathered counts; see the module docstring."""
    from . import _lib as L
    lib = engine.lib
    # a chromosome without observed regions has no handles (calculatePvalues returned NULLs, R/calculatePvalues.R:
    # 245-251): it takes part in the collectives with an empty record
    R = int(lib.dfb_regions_count(regions_handle)) if regions_handle else 0
    marks = []

    def mark(name):
        if os.environ.get("DFB_POOL_PROFILE"):
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            marks.append((name, ev))

    mark("start")
    active = dist is not None and dist.is_available() and dist.is_initialized()
    world = dist.get_world_size() if active else 1
    rank = dist.get_rank() if active else 0
    key = (device.index, id(dist) if active else None)
    if key not in _capacity:  # first call: agree on a capacity (side stream: the main stream keeps running)
        if key not in _side_streams:
            _side_streams[key] = torch.cuda.Stream(device=device)
        counts = exchange_counts([R], device, dist=dist if active else None, side_stream=_side_streams[key])
        _capacity[key] = _headroom(max(c[0] for c in counts))
    # the engine and torch share one stream in bench.py; otherwise order them explicitly
    shared = engine.stream_handle == torch.cuda.current_stream().cuda_stream

    def to_engine():
        if not shared:
            torch.cuda.current_stream().synchronize()

    def to_torch():
        if not shared:
            engine.sync()

    ma = 1 if merge_area else 0
    p_dev = torch.empty(max(R, 1), device=device, dtype=torch.float64)
    fwer = torch.empty(max(R, 1), device=device, dtype=torch.float64)
    while True:
        rcap = _capacity[key]
        record = torch.empty(2 + B + rcap, device=device, dtype=torch.float64)
        merged = torch.empty(world * rcap, device=device, dtype=torch.float64)
        src = torch.empty(world * rcap, device=device, dtype=torch.int32)
        hist = torch.empty(world * rcap + 1, device=device, dtype=torch.int64)
        to_engine()
        L.check(lib.dfb_pool_local_dev(engine.h, regions_handle, nulls_handle, B, ma, rcap,
                                       C.c_void_p(record.data_ptr())))
        to_torch()
        mark("local")
        records = gather_records(record, dist=dist if active else None)
        mark("gather")
        to_engine()
        L.check(lib.dfb_pool_hist_dev(engine.h, nulls_handle, ma, C.c_void_p(records.data_ptr()), world, B, rcap,
                                      C.c_void_p(merged.data_ptr()), C.c_void_p(src.data_ptr()),
                                      C.c_void_p(hist.data_ptr())))
        to_torch()
        mark("hist")
        if active:
            reduce_histogram(hist, dist=dist)
        mark("reduce")
        to_engine()
        L.check(lib.dfb_pool_rank_dev(engine.h, regions_handle, C.c_void_p(records.data_ptr()), world, rank, B, rcap,
                                      C.c_void_p(merged.data_ptr()), C.c_void_p(src.data_ptr()),
                                      C.c_void_p(hist.data_ptr()), float(cutoff_used), C.c_void_p(p_dev.data_ptr()),
                                      C.c_void_p(fwer.data_ptr())))
        to_torch()
        mark("rank")
        head = records[:, :2].to("cpu")  # everything is queued: the one host read (identical on every rank)
        biggest = int(head[:, 0].max().item())
        if biggest <= rcap:
            break
        _capacity[key] = _headroom(biggest)  # somebody outgrew the records: every rank repeats
    if marks:
        PROFILE["calls"] += 1
        PROFILE["events"].append(marks)
    return p_dev[:R], fwer[:R], int(head[:, 1].sum().item())


# ---------------------------------------------------------------------------------------------
# one-call mergeResults on the library's own NCCL communicator (no torch on the data path)


def init_communicator(engine, dist=None, device=None):
    """Give `engine` its rank in the job's NCCL communicator.  The library creates the communicator
    itself (dfb_comm_unique_id on rank 0, dfb_comm_init_rank everywhere); the process group is only the
    courier of the 128-byte id -- any host-side channel would do (the R glue uses a file)."""
    import numpy as np
    from . import _lib as L
    if dist is None:
        import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return 1, 0
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    buf = (C.c_ubyte * 128)()
    if rank == 0:
        L.check(engine.lib.dfb_comm_unique_id(buf))
    t = torch.tensor(list(bytes(buf)), dtype=torch.uint8)
    if dist.get_backend() == "nccl":
        t = t.to(device if device is not None else torch.device("cuda", engine.device))
    dist.broadcast(t, 0)
    raw = bytes(np.asarray(t.cpu().numpy(), dtype=np.uint8).tobytes())
    idb = (C.c_ubyte * 128).from_buffer_copy(raw)
    L.check(engine.lib.dfb_comm_init_rank(engine.h, world, rank, idb))
    return world, rank


def merge_results(engine, chromosomes, B, cutoff_used, merge_area=True, want=True):
    """mergeResults() for the chromosomes this rank holds (R/mergeResults.R:216-235, 300-307, 380-386) as
    ONE library call; collective over the engine's communicator.  `chromosomes`: list of
    (regions_handle, nulls_handle), (None, None) for a chromosome without observed regions.  Returns
    ([(pvalues, fwer) per chromosome, regions in decreasing-area order], genome-wide null count)."""
    import numpy as np
    from . import _lib as L
    lib = engine.lib
    k = len(chromosomes)
    rp = (C.c_void_p * max(k, 1))(*[(rh.value if isinstance(rh, C.c_void_p) else rh) for rh, _ in chromosomes])
    npp = (C.c_void_p * max(k, 1))(*[(nh.value if isinstance(nh, C.c_void_p) else nh) for _, nh in chromosomes])
    outs = []
    pp = (C.c_void_p * max(k, 1))()
    fp = (C.c_void_p * max(k, 1))()
    for c, (rh, _) in enumerate(chromosomes):
        R = int(lib.dfb_regions_count(rh)) if rh else 0
        pv, fw = np.empty(R), np.empty(R)
        outs.append((pv, fw))
        pp[c] = pv.ctypes.data if (want and R) else None
        fp[c] = fw.ctypes.data if (want and R) else None
    total = C.c_int64()
    L.check(lib.dfb_merge_results(engine.h, k, rp, npp, int(B), float(cutoff_used), 1 if merge_area else 0, pp, fp,
                                  C.byref(total)))
    return outs, int(total.value)


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
