// derfinder_b200 -- shared internals of the CUDA library (context, error plumbing, device buffers,
// kernel timing).  sm_100a only; there is no CPU path in this library.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/derfinder_b200.h"

namespace dfb {

void set_error(const char* fmt, ...);

#define DFB_CUDA(expr)                                                                         \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            dfb::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return DFB_ECUDA;                                                                  \
        }                                                                                      \
    } while (0)

#define DFB_TRY(expr)             \
    do {                          \
        int _rc = (expr);         \
        if (_rc < 0) return _rc;  \
    } while (0)

#define DFB_REQUIRE(cond, ...)          \
    do {                                \
        if (!(cond)) {                  \
            dfb::set_error(__VA_ARGS__); \
            return DFB_EINVAL;          \
        }                               \
    } while (0)

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t round_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

}  // namespace dfb

namespace dfb {

// Small persistent host thread pool used only to move bytes between pinned staging buffers and the
// caller's pageable memory at more than one core's memcpy bandwidth (PCIe 5 x16 outruns a single
// core).  parallel_for(n, fn) runs fn(0..n-1) on the workers plus the calling thread.
class CopyPool {
  public:
    explicit CopyPool(int nthreads);
    ~CopyPool();
    int size() const { return (int)workers_.size() + 1; }
    void parallel_for(int nparts, const std::function<void(int)>& fn);

  private:
    void worker();
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable work_cv_, done_cv_;
    const std::function<void(int)>* fn_ = nullptr;
    std::atomic<int> next_{0};
    std::atomic<int> active_{0};
    std::atomic<uint64_t> generation_{0};
    int nparts_ = 0;
    std::atomic<bool> stop_{false};
};

}  // namespace dfb

struct dfb_ctx {
    int device = 0;
    cudaMemPool_t mempool = nullptr;  // private stream-ordered allocation pool of this context
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sm_count = 148;
    size_t smem_optin = 0;
    int64_t launches = 0;
    unsigned timing = 0;  // bit k: launches of kernel family k are bracketed by CUDA events
    std::vector<cudaEvent_t> ev_free;  // recycled timing events (creating two per launch costs more than the launch)
    struct Ev {
        cudaEvent_t a, b;
        int k;
        const char* file;
        int line;
    };
    std::vector<Ev> evs;
    double kms[DFB_K_COUNT] = {0};
    int64_t klaunch[DFB_K_COUNT] = {0};
    // options
    int opt_screen = 1;
    int opt_dbg = 0;             // developer timing experiments (results invalid when non-zero)
    int opt_epi_warps = 0;       // fused kernel: epilogue warps per TMEM quarter (0 = auto)
    int opt_b_stages = 0;        // fused kernel: B tile stages (0 = auto)
    int opt_ys_smem = -1;        // fused kernel: FP64 tile in shared memory (1), streamed (0), auto (-1)
    int opt_chunk_perms = 0;     // fused kernel: permutations per MMA chunk (0 = auto)
    int opt_tmem_stages = 0;     // fused kernel: accumulator stages (0 = auto)
    int opt_fused = 1;           // fused screen+refine kernel (0: two-kernel screen path)
    int opt_a16_fused = 1;       // observed-F kernel also writes the null scan's fp16 operand (0: built on first use)
    int opt_a_stages = 0;        // null2 kernel: A tile ring depth (0 = auto)
    int opt_mma_warps = 0;       // null2 kernel: MMA-issuing warps (0 = auto)
    int opt_lazy_y = 1;          // integer coverage: log2 matrix not materialised by dfb_preprocess (see dfb_cov::y_lazy)
    int opt_null2 = 1;           // second-generation null scan (fstat_tc2.cu): 1 = when it pays (see null2_wanted), 2 = always, 0 = never
    int opt_k2_waves = 16;       // fused kernel: CTAs per resident slot (1 = persistent); see launch_fused
    double opt_null_cap_mb = 0;  // 0 = auto
    int opt_perm_batch = 0;      // 0 = auto
    // Mask rows of the tensor-core null scan, kept between calls and kept ZERO: the scan writes its non-empty
    // words only, the null-region stage reads them, and a clean-up kernel zeroes exactly the groups that were
    // flagged -- instead of a memset of rows * W words (3 GB at the chr1 shape) before every scan.
    uint32_t* k2_mask = nullptr;
    size_t k2_mask_cap = 0;        // words
    unsigned char* k2_gflag = nullptr;
    size_t k2_gflag_cap = 0;       // bytes
    bool k2_clean = false;         // both buffers are all zero
    int opt_mask_cache = 1;
    int opt_cluster_one_block = 1;  // cluster chain of small region tables by one block in one launch
    int opt_pinned_direct = 1;   // Rle columns in page-locked caller memory are DMAed in place (dfb_cov_from_rle)
    int opt_poison = 0;          // DFB_POISON=1: fill every fresh device allocation with 0xff (uninitialised-read hunt)
    int opt_wire16 = 1;          // short-run integer Rle columns cross PCIe as 16-bit values / lengths (see upload_rle)
    int opt_filter_events = 1;   // filterData mask of integer coverage from run events (0: expansion kernel)
    double null_density = 0;     // exceedance values per (base, permutation) seen by the last null scan (0 = unknown)
    double opt_scratch_gb = 24;  // upper bound for per-call scratch (mask + exceedance values)
    // pinned staging ring for pageable <-> device copies, and the host copy pool that drains it
    void* pinned = nullptr;
    size_t pinned_bytes = 0;
    cudaEvent_t stage_ev[4] = {nullptr, nullptr, nullptr, nullptr};
    // h2d_gather pipelines the same ring in 2 MiB sub-stages (a 10 MB Rle upload overlaps its packing with
    // its DMA instead of alternating 8 MiB of each); the counter runs on across calls
    std::vector<cudaEvent_t> gather_ev;
    size_t gather_chunk = 0;
    dfb::CopyPool* pool = nullptr;
    // pinned arena for scalar / O(regions) read-backs: D2H into pinned memory is truly asynchronous,
    // so several results can be fetched with ONE stream synchronisation.  Blocks are kept for reuse.
    struct PinBlock {
        char* p;
        size_t cap, used;
    };
    std::vector<PinBlock> pin_blocks;
    // auxiliary stream: the observed-region chain (dozens of tiny launches) runs here while the
    // permutation scan occupies the main stream
    cudaStream_t aux = nullptr;
    cudaEvent_t ev_main = nullptr, ev_aux = nullptr;
    // background download (dfb_fstats_begin): its own stream, two pinned stages and one host thread, so a
    // large result crosses PCIe and lands in the caller's pageable vector while the main stream computes
    struct Background {
        cudaStream_t stream = nullptr;
        cudaEvent_t ready = nullptr, ev[2] = {nullptr, nullptr};
        char* pin = nullptr;
        std::thread th;
        bool active = false;
        int rc = 0;
    } bg;
    // multi-GPU exchange (comm.cu): NCCL communicator of this context's rank, null for a single GPU
    void* nccl_comm = nullptr;
    int world = 1, rank = 0;
    // log2(k + scalefac) table for integer coverage (glibc log2, built once per scalefac)
    double* lut_dev = nullptr;
    int64_t lut_n = 0;
    double lut_scalefac = 0;
};

namespace dfb {

// Brackets the launches of one kernel family with CUDA events (when timing is on) and counts
// launches.  Usage: { KTimer t(ctx, DFB_K_FSTAT, nlaunches); kernel<<<...>>>(); }
struct KTimer {
    dfb_ctx* c;
    int k;
    cudaEvent_t a = nullptr, b = nullptr;
    const char* file;
    int line;
    KTimer(dfb_ctx* ctx, int kernel, int nlaunch = 1, const char* f = __builtin_FILE(), int l = __builtin_LINE())
        : c(ctx), k(kernel), file(f), line(l) {
        c->launches += nlaunch;
        c->klaunch[k] += nlaunch;
        if ((c->timing >> k) & 1u) {
            a = take(c);
            b = take(c);
            cudaEventRecord(a, c->stream);
        }
    }
    ~KTimer() {
        if (a) {
            cudaEventRecord(b, c->stream);
            c->evs.push_back({a, b, k, file, line});
        }
    }
    static cudaEvent_t take(dfb_ctx* c) {
        cudaEvent_t e = nullptr;
        if (!c->ev_free.empty()) {
            e = c->ev_free.back();
            c->ev_free.pop_back();
        } else {
            cudaEventCreate(&e);
        }
        return e;
    }
};

typedef cudaStream_t dfb_ctx_stream_t;
// Stream-ordered device buffer (cudaMallocAsync pool: repeated calls reuse memory without
// device-wide synchronisation).
template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    cudaStream_t s = nullptr;
    bool borrowed = false;  // p belongs to the caller (dfb_cov_view_dense_dev): never freed here
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n), s(o.s), borrowed(o.borrowed) { o.p = nullptr; o.n = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) {
            release();
            p = o.p; n = o.n; s = o.s; borrowed = o.borrowed;
            o.p = nullptr; o.n = 0;
        }
        return *this;
    }
    ~DevBuf() { release(); }
    void release() {
        if (p && !borrowed) cudaFreeAsync(p, s);
        p = nullptr;
        n = 0;
        borrowed = false;
    }
    void borrow(dfb_ctx_stream_t stream, T* ptr, size_t count) {
        release();
        p = ptr;
        n = count;
        s = stream;
        borrowed = true;
    }
    int alloc(dfb_ctx* ctx, size_t count) {
        release();
        s = ctx->stream;
        n = count;
        if (count == 0) return DFB_OK;
        cudaError_t e = ctx->mempool ? cudaMallocFromPoolAsync((void**)&p, count * sizeof(T), ctx->mempool, s)
                                     : cudaMallocAsync((void**)&p, count * sizeof(T), s);
        if (e != cudaSuccess) {
            p = nullptr;
            n = 0;
            set_error("device allocation of %zu bytes failed: %s", count * sizeof(T), cudaGetErrorString(e));
            cudaGetLastError();
            return DFB_ENOMEM;
        }
        // developer check (DFB_POISON=1): fresh allocations are filled with 0xff (NaN / -1 / huge) so that a read
        // of memory nobody wrote shows up in the parity tests instead of hiding behind recycled pool contents
        if (ctx->opt_poison) cudaMemsetAsync(p, 0xff, count * sizeof(T), s);
        return DFB_OK;
    }
    int zero() {
        if (p && n) {
            DFB_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s));
        }
        return DFB_OK;
    }
};

// bump allocation from the context's pinned arena (valid until the next arena_reset)
void* arena_alloc(dfb_ctx* ctx, size_t bytes);
void arena_reset(dfb_ctx* ctx);
// give back the most recent arena_alloc of a block (no-op if something was allocated after it)
void arena_release(dfb_ctx* ctx, void* ptr, size_t bytes);
// asynchronous D2H into arena memory: returns the pinned pointer, valid after the next stream sync
template <typename T>
static inline T* d2h_async_pinned(dfb_ctx* ctx, const T* src_dev, size_t count) {
    T* h = (T*)arena_alloc(ctx, std::max<size_t>(count, 1) * sizeof(T));
    if (h && count) {
        if (cudaMemcpyAsync(h, src_dev, count * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) return nullptr;
    }
    return h;
}
int h2d(dfb_ctx* ctx, void* dst, const void* src, size_t bytes);
int d2h(dfb_ctx* ctx, void* dst, const void* src, size_t bytes);  // synchronises the stream
// pageable <-> device through a ring of pinned staging buffers drained / filled by the host copy
// pool; both return with the copy complete.  A source/destination that is already pinned (or
// registered) host memory is copied directly.
int d2h_big(dfb_ctx* ctx, void* dst, const void* src, size_t bytes);
// the same copy in the background, ordered after the work queued on the stream so far; dst must not
// be touched until bg_download_wait (every other pooled transfer of the context waits for it too)
int bg_download_begin(dfb_ctx* ctx, void* dst, const void* src, size_t bytes);
int bg_download_wait(dfb_ctx* ctx);
int h2d_big(dfb_ctx* ctx, void* dst, const void* src, size_t bytes);
// H2D of `count` host segments packed back to back at dst (through the pinned ring); synchronises.
int h2d_gather(dfb_ctx* ctx, void* dst, const void* const* srcs, const size_t* bytes, int count, bool sync = true);
int ensure_aux_stream(dfb_ctx* ctx);
int copy_threads(dfb_ctx* ctx);
bool host_is_pinned(const void* p);  // page-locked (cudaHostAlloc / cudaHostRegister) host memory
int h2d_gather_narrow16(dfb_ctx* ctx, uint16_t* dst, const int32_t* const* srcs, const int64_t* counts, int count, bool sync,
                        bool* overflow);
// D2H of `count` elements of `elem` bytes laid out in B consecutive blocks (block b = elements
// [pre[b], pre[b+1])) into the reference's duplicated layout: block b is written twice, back to
// back, at dst element 2*pre[b] (R/calculatePvalues.R:346-354).
int d2h_dup(dfb_ctx* ctx, void* dst, const void* src, size_t elem, const int64_t* pre, int64_t B);
// several arrays with the same block structure as one DMA / copy pipeline
int d2h_dup_multi(dfb_ctx* ctx, int na, void* const* dsts, const void* const* srcs, const size_t* elems,
                  const int64_t* pre, int64_t B);

// device-wide exclusive scan of `n` uint32 counts into int64 offsets; total_dev[0] = sum.
int exclusive_scan_u32(dfb_ctx* ctx, const uint32_t* in, int64_t* out, int64_t n, int64_t* total_dev,
                       int kernel_family);

// comm.cu: collectives on the context's stream (copies / no-ops for a single GPU)
int comm_all_gather(dfb_ctx* ctx, const void* send, void* recv, size_t count, int dtype_bytes);
int comm_all_reduce_sum_u64(dfb_ctx* ctx, void* buf, size_t count);

static inline int grid_for(int64_t work_items, int threads, int sm_count, int max_waves = 32) {
    int64_t blocks = ceil_div(work_items, threads);
    int64_t cap = (int64_t)sm_count * max_waves;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

}  // namespace dfb
