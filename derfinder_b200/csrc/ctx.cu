// Context, error string, copies, device-wide scan.
#include <emmintrin.h>

#include <chrono>
#include <stdarg.h>
#include <stdlib.h>

#include <algorithm>
#include <atomic>

#include "common.cuh"

namespace dfb {

static thread_local std::string g_err;

void set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
}

void* arena_alloc(dfb_ctx* ctx, size_t bytes) {
    bytes = (bytes + 63) & ~(size_t)63;
    for (auto& b : ctx->pin_blocks)
        if (b.cap - b.used >= bytes) {
            void* p = b.p + b.used;
            b.used += bytes;
            return p;
        }
    const size_t cap = std::max<size_t>(bytes, (size_t)4 << 20);
    char* p = nullptr;
    if (cudaMallocHost((void**)&p, cap) != cudaSuccess) {
        cudaGetLastError();
        set_error("pinned host allocation of %zu bytes failed", cap);
        return nullptr;
    }
    ctx->pin_blocks.push_back({p, cap, bytes});
    return p;
}

void arena_release(dfb_ctx* ctx, void* ptr, size_t bytes) {
    bytes = (bytes + 63) & ~(size_t)63;
    for (auto& b : ctx->pin_blocks)
        if (b.p + b.used == (char*)ptr + bytes && b.used >= bytes) {
            b.used -= bytes;
            return;
        }
}

void arena_reset(dfb_ctx* ctx) {
    for (auto& b : ctx->pin_blocks) b.used = 0;
}

int h2d(dfb_ctx* ctx, void* dst, const void* src, size_t bytes) {
    if (bytes == 0) return DFB_OK;
    DFB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return DFB_OK;
}

int d2h(dfb_ctx* ctx, void* dst, const void* src, size_t bytes) {
    if (bytes) DFB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    DFB_CUDA(cudaStreamSynchronize(ctx->stream));
    return DFB_OK;
}

// ------------------------------------------------------------------------------------------
// host copy pool

CopyPool::CopyPool(int nthreads) {
    for (int i = 0; i < nthreads - 1; ++i) workers_.emplace_back([this] { worker(); });
}

CopyPool::~CopyPool() {
    {
        std::lock_guard<std::mutex> lk(m_);
        stop_ = true;
        generation_.fetch_add(1, std::memory_order_release);  // wake spinners too
    }
    work_cv_.notify_all();
    for (auto& t : workers_) t.join();
}

// Workers spin for a short while after a job before they block: the chunks of one transfer arrive
// every ~150 us, and a futex wake-up (hundreds of us on a busy VM) would cost more than the copy.
static inline void cpu_relax() { _mm_pause(); }

void CopyPool::worker() {
    uint64_t seen = 0;
    for (;;) {
        bool have = false;
        const auto t0 = std::chrono::steady_clock::now();
        for (unsigned spin = 0;; ++spin) {  // stay hot for a few ms: one pipeline step issues several transfers
            if (generation_.load(std::memory_order_acquire) != seen) {
                have = true;
                break;
            }
            cpu_relax();
            if ((spin & 255) == 255 &&
                std::chrono::steady_clock::now() - t0 > std::chrono::milliseconds(6)) break;
        }
        if (!have) {
            std::unique_lock<std::mutex> lk(m_);
            work_cv_.wait(lk, [&] { return stop_.load() || generation_.load(std::memory_order_acquire) != seen; });
        }
        if (stop_.load()) return;
        seen = generation_.load(std::memory_order_acquire);
        const std::function<void(int)>* fn = fn_;
        const int nparts = nparts_;
        for (int i; (i = next_.fetch_add(1)) < nparts;) (*fn)(i);
        if (active_.fetch_sub(1, std::memory_order_acq_rel) == 1) {
            std::lock_guard<std::mutex> lk(m_);
            done_cv_.notify_all();
        }
    }
}

void CopyPool::parallel_for(int nparts, const std::function<void(int)>& fn) {
    if (nparts <= 1 || workers_.empty()) {
        for (int i = 0; i < nparts; ++i) fn(i);
        return;
    }
    {
        std::lock_guard<std::mutex> lk(m_);
        fn_ = &fn;
        nparts_ = nparts;
        next_.store(0);
        active_.store((int)workers_.size());
        generation_.fetch_add(1, std::memory_order_release);
    }
    work_cv_.notify_all();
    for (int i; (i = next_.fetch_add(1)) < nparts;) fn(i);
    for (int spin = 0; spin < 2000000; ++spin) {
        if (active_.load(std::memory_order_acquire) == 0) return;
        cpu_relax();
    }
    std::unique_lock<std::mutex> lk(m_);
    done_cv_.wait(lk, [&] { return active_.load(std::memory_order_acquire) == 0; });
}

// Large pageable host buffers (R vectors, numpy arrays) are moved through a ring of pinned staging
// buffers so the DMA runs at PCIe speed; the host side of every chunk (pinned <-> pageable memcpy)
// is spread over the copy pool and overlaps the DMA of the following chunks.
constexpr size_t STAGE_BYTES = 8u << 20;
constexpr int NSTAGE = 4;

static int ensure_staging(dfb_ctx* ctx) {
    if (ctx->bg.active) DFB_TRY(bg_download_wait(ctx));  // the copy pool serves one transfer at a time
    if (ctx->pinned) return DFB_OK;
    DFB_CUDA(cudaMallocHost(&ctx->pinned, NSTAGE * STAGE_BYTES));
    ctx->pinned_bytes = NSTAGE * STAGE_BYTES;
    for (int i = 0; i < NSTAGE; ++i) DFB_CUDA(cudaEventCreateWithFlags(&ctx->stage_ev[i], cudaEventDisableTiming));
    if (!ctx->pool) {
        // three quarters of the cores, at most 12 -- shared between the processes of one node (one per GPU): the
        // launcher's LOCAL_WORLD_SIZE tells how many there are.  Spinning workers of 8 ranks on a
        // 16-core host would otherwise starve each other.
        unsigned hc = std::thread::hardware_concurrency();
        unsigned share = 1;
        if (const char* e = getenv("LOCAL_WORLD_SIZE")) share = (unsigned)std::max(1, atoi(e));
        // measured on the 16-core bench host: 12 threads beat 8 on every transfer, 16 start to starve the
        // thread that drives the GPU
        int nt = (int)std::min<unsigned>(12, std::max<unsigned>(1, (hc * 3 / 4) / share));
        if (const char* e = getenv("DFB_COPY_THREADS")) nt = std::max(1, std::min(64, atoi(e)));
        ctx->pool = new CopyPool(nt);
    }
    return DFB_OK;
}

// The auxiliary stream (+ the two events that order it against the main one), created on first use.  Highest
// priority: when the null scan's CTAs retire, the freed SM slots go to the observed-region chain first.
int ensure_aux_stream(dfb_ctx* ctx) {
    if (ctx->aux) return DFB_OK;
    int prio_lo = 0, prio_hi = 0;
    DFB_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    DFB_CUDA(cudaStreamCreateWithPriority(&ctx->aux, cudaStreamNonBlocking, prio_hi));
    DFB_CUDA(cudaEventCreateWithFlags(&ctx->ev_main, cudaEventDisableTiming));
    DFB_CUDA(cudaEventCreateWithFlags(&ctx->ev_aux, cudaEventDisableTiming));
    return DFB_OK;
}

// threads of the host copy pool (created on first use)
int copy_threads(dfb_ctx* ctx) {
    if (ensure_staging(ctx) < 0 || !ctx->pool) return 1;
    return ctx->pool->size();
}

constexpr size_t SUB_BYTES = 2u << 20;                       // sub-stage of the ring
constexpr size_t NSUB = NSTAGE * STAGE_BYTES / SUB_BYTES;      // sub-stages in the ring
static int ensure_substage_events(dfb_ctx* ctx) {
    if (!ctx->gather_ev.empty()) return DFB_OK;
    ctx->gather_ev.resize(NSUB, nullptr);
    for (auto& e : ctx->gather_ev) DFB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    return DFB_OK;
}

bool host_is_pinned(const void* p);
static bool is_pinned_host(const void* p) { return host_is_pinned(p); }
bool host_is_pinned(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

// memcpy with non-temporal stores: the destination (the caller's result vector) is written once and
// not read back here, so bypassing the cache avoids the read-for-ownership traffic of every line.
static void copy_stream(char* dst, const char* src, size_t n) {
    if (n < 4096) {
        memcpy(dst, src, n);
        return;
    }
    const size_t head = (16 - ((uintptr_t)dst & 15)) & 15;
    memcpy(dst, src, head);
    dst += head;
    src += head;
    n -= head;
    const size_t body = n & ~(size_t)63;
    for (size_t i = 0; i < body; i += 64) {
        const __m128i a = _mm_loadu_si128((const __m128i*)(src + i));
        const __m128i b = _mm_loadu_si128((const __m128i*)(src + i + 16));
        const __m128i c = _mm_loadu_si128((const __m128i*)(src + i + 32));
        const __m128i d = _mm_loadu_si128((const __m128i*)(src + i + 48));
        _mm_stream_si128((__m128i*)(dst + i), a);
        _mm_stream_si128((__m128i*)(dst + i + 16), b);
        _mm_stream_si128((__m128i*)(dst + i + 32), c);
        _mm_stream_si128((__m128i*)(dst + i + 48), d);
    }
    memcpy(dst + body, src + body, n - body);
    _mm_sfence();
}

static void pool_memcpy(dfb_ctx* ctx, char* dst, const char* src, size_t bytes) {
    const int parts = (int)std::min<size_t>((size_t)ctx->pool->size(), std::max<size_t>(1, bytes >> 18));
    const size_t per = ((bytes + parts - 1) / parts + 63) & ~(size_t)63;
    ctx->pool->parallel_for(parts, [&](int i) {
        const size_t o = (size_t)i * per;
        if (o < bytes) copy_stream(dst + o, src + o, std::min(per, bytes - o));
    });
}

// Generic chunked D2H: `sink(chunk_offset, pinned_ptr, chunk_bytes)` consumes each chunk on the host.
template <typename Sink>
static int d2h_chunks(dfb_ctx* ctx, const void* src, size_t bytes, Sink sink) {
    DFB_TRY(ensure_staging(ctx));
    char* pin = (char*)ctx->pinned;
    const size_t nchunk = (bytes + STAGE_BYTES - 1) / STAGE_BYTES;
    size_t issued = 0;
    for (size_t done = 0; done < nchunk; ++done) {
        for (; issued < nchunk && issued < done + NSTAGE; ++issued) {
            const size_t off = issued * STAGE_BYTES, sz = std::min(STAGE_BYTES, bytes - off);
            DFB_CUDA(cudaMemcpyAsync(pin + (issued % NSTAGE) * STAGE_BYTES, (const char*)src + off, sz,
                                     cudaMemcpyDeviceToHost, ctx->stream));
            DFB_CUDA(cudaEventRecord(ctx->stage_ev[issued % NSTAGE], ctx->stream));
        }
        const size_t off = done * STAGE_BYTES, sz = std::min(STAGE_BYTES, bytes - off);
        DFB_CUDA(cudaEventSynchronize(ctx->stage_ev[done % NSTAGE]));
        sink(off, pin + (done % NSTAGE) * STAGE_BYTES, sz);
    }
    return DFB_OK;
}

int bg_download_wait(dfb_ctx* ctx) {
    if (!ctx->bg.active) return DFB_OK;
    ctx->bg.th.join();
    ctx->bg.active = false;
    if (ctx->bg.rc < 0) {
        set_error("background download failed");
        return ctx->bg.rc;
    }
    return DFB_OK;
}

int bg_download_begin(dfb_ctx* ctx, void* dst, const void* src, size_t bytes) {
    DFB_TRY(ensure_staging(ctx));  // waits for a previous background download; creates the copy pool
    if (bytes == 0) return DFB_OK;
    auto& g = ctx->bg;
    if (!g.stream) {
        DFB_CUDA(cudaStreamCreateWithFlags(&g.stream, cudaStreamNonBlocking));
        DFB_CUDA(cudaEventCreateWithFlags(&g.ready, cudaEventDisableTiming));
        for (int i = 0; i < 2; ++i) DFB_CUDA(cudaEventCreateWithFlags(&g.ev[i], cudaEventDisableTiming));
        DFB_CUDA(cudaMallocHost((void**)&g.pin, 2 * STAGE_BYTES));
    }
    DFB_CUDA(cudaEventRecord(g.ready, ctx->stream));
    DFB_CUDA(cudaStreamWaitEvent(g.stream, g.ready, 0));
    g.rc = DFB_OK;
    g.active = true;
    g.th = std::thread([ctx, dst, src, bytes] {
        auto& b = ctx->bg;
        if (cudaSetDevice(ctx->device) != cudaSuccess) {
            b.rc = DFB_ECUDA;
            return;
        }
        const size_t nchunk = (bytes + STAGE_BYTES - 1) / STAGE_BYTES;
        size_t issued = 0;
        for (size_t done = 0; done < nchunk; ++done) {
            for (; issued < nchunk && issued < done + 2; ++issued) {
                const size_t off = issued * STAGE_BYTES, sz = std::min(STAGE_BYTES, bytes - off);
                if (cudaMemcpyAsync(b.pin + (issued & 1) * STAGE_BYTES, (const char*)src + off, sz, cudaMemcpyDeviceToHost,
                                    b.stream) != cudaSuccess ||
                    cudaEventRecord(b.ev[issued & 1], b.stream) != cudaSuccess) {
                    b.rc = DFB_ECUDA;
                    return;
                }
            }
            if (cudaEventSynchronize(b.ev[done & 1]) != cudaSuccess) {
                b.rc = DFB_ECUDA;
                return;
            }
            const size_t off = done * STAGE_BYTES, sz = std::min(STAGE_BYTES, bytes - off);
            pool_memcpy(ctx, (char*)dst + off, b.pin + (done & 1) * STAGE_BYTES, sz);
        }
    });
    return DFB_OK;
}

int d2h_big(dfb_ctx* ctx, void* dst, const void* src, size_t bytes) {
    if (bytes < (1u << 20) || is_pinned_host(dst)) return d2h(ctx, dst, src, bytes);
    return d2h_chunks(ctx, src, bytes,
                      [&](size_t off, const char* p, size_t sz) { pool_memcpy(ctx, (char*)dst + off, p, sz); });
}

// Chunked D2H of several device arrays as ONE pipeline: the DMA of the next array's first chunk is in
// flight while the host still drains the previous array.  sink(array, chunk_offset, pinned, bytes).
template <typename Sink>
static int d2h_chunks_multi(dfb_ctx* ctx, int na, const void* const* srcs, const size_t* bytes, Sink sink) {
    DFB_TRY(ensure_staging(ctx));
    DFB_TRY(ensure_substage_events(ctx));
    // 2 MiB sub-stages of the ring, 16 in flight: the whole ~30 MB of a C2 result is requested up front
    // and the host drains each piece as it lands (8 MiB stages left 0.2 ms of DMA and 0.2 ms of host
    // copy un-overlapped at the two ends)
    char* pin = (char*)ctx->pinned;
    struct Chunk { int a; size_t off, sz; };
    std::vector<Chunk> chunks;
    for (int a = 0; a < na; ++a)
        for (size_t off = 0; off < bytes[a]; off += SUB_BYTES) chunks.push_back({a, off, std::min(SUB_BYTES, bytes[a] - off)});
    size_t issued = 0;
    for (size_t done = 0; done < chunks.size(); ++done) {
        for (; issued < chunks.size() && issued < done + NSUB; ++issued) {
            const Chunk& c = chunks[issued];
            DFB_CUDA(cudaMemcpyAsync(pin + (issued % NSUB) * SUB_BYTES, (const char*)srcs[c.a] + c.off, c.sz,
                                     cudaMemcpyDeviceToHost, ctx->stream));
            DFB_CUDA(cudaEventRecord(ctx->gather_ev[issued % NSUB], ctx->stream));
        }
        DFB_CUDA(cudaEventSynchronize(ctx->gather_ev[done % NSUB]));
        const Chunk& c = chunks[done];
        sink(c.a, c.off, pin + (done % NSUB) * SUB_BYTES, c.sz);
    }
    return DFB_OK;
}

int d2h_dup_multi(dfb_ctx* ctx, int na, void* const* dsts, const void* const* srcs, const size_t* elems,
                  const int64_t* pre, int64_t B) {
    const size_t count = (size_t)pre[B];
    if (count == 0 || na <= 0) return DFB_OK;
    std::vector<size_t> bytes((size_t)na);
    std::vector<int64_t> blk((size_t)na, 0);  // first block that may overlap the array's current chunk
    for (int a = 0; a < na; ++a) bytes[(size_t)a] = count * elems[a];
    return d2h_chunks_multi(ctx, na, srcs, bytes.data(), [&](int a, size_t off, const char* p, size_t sz) {
        // element range of this chunk (STAGE_BYTES is a multiple of every element size used)
        const size_t elem = elems[a];
        const int64_t e0 = (int64_t)(off / elem), e1 = e0 + (int64_t)(sz / elem);
        struct Seg { char* d1; char* d2; const char* s; size_t n; };
        std::vector<Seg> parts;
        int64_t& bk = blk[(size_t)a];
        while (bk < B && pre[bk + 1] <= e0) ++bk;
        for (int64_t b = bk; b < B && pre[b] < e1; ++b) {
            const int64_t lo = std::max(pre[b], e0), hi = std::min(pre[b + 1], e1);
            if (hi <= lo) continue;
            const int64_t k = pre[b + 1] - pre[b];
            char* d1 = (char*)dsts[a] + (size_t)(2 * pre[b] + (lo - pre[b])) * elem;
            const char* sp = p + (size_t)(lo - e0) * elem;
            const size_t nb = (size_t)(hi - lo) * elem;
            for (size_t o = 0; o < nb; o += (128u << 10))  // ~128 KiB pieces for the pool
                parts.push_back({d1 + o, d1 + (size_t)k * elem + o, sp + o, std::min<size_t>(128u << 10, nb - o)});
        }
        ctx->pool->parallel_for((int)parts.size(), [&](int i) {
            copy_stream(parts[i].d1, parts[i].s, parts[i].n);
            copy_stream(parts[i].d2, parts[i].s, parts[i].n);
        });
    });
}

int d2h_dup(dfb_ctx* ctx, void* dst, const void* src, size_t elem, const int64_t* pre, int64_t B) {
    void* const dsts[1] = {dst};
    const void* const srcs[1] = {src};
    return d2h_dup_multi(ctx, 1, dsts, srcs, &elem, pre, B);
}

// H2D of `count` host segments laid back to back in device memory: the segments are packed into the
// pinned ring by the copy pool (one DMA per 8 MiB instead of one driver-staged copy per segment).
int h2d_gather(dfb_ctx* ctx, void* dst, const void* const* srcs, const size_t* bytes, int count, bool sync) {
    DFB_TRY(ensure_staging(ctx));
    constexpr size_t GS = SUB_BYTES, NG = NSUB;
    DFB_TRY(ensure_substage_events(ctx));
    char* pin = (char*)ctx->pinned;
    struct Piece { char* d; const char* s; size_t n; };
    size_t& chunk = ctx->gather_chunk;  // persistent: a call that does not synchronise leaves DMAs in flight
    size_t fill = 0, dev_off = 0;
    std::vector<Piece> pieces;
    bool waited = false;
    auto flush = [&]() -> int {
        if (fill == 0) return DFB_OK;
        ctx->pool->parallel_for((int)pieces.size(), [&](int i) { copy_stream(pieces[i].d, pieces[i].s, pieces[i].n); });
        DFB_CUDA(cudaMemcpyAsync((char*)dst + dev_off, pin + (chunk % NG) * GS, fill, cudaMemcpyHostToDevice, ctx->stream));
        DFB_CUDA(cudaEventRecord(ctx->gather_ev[chunk % NG], ctx->stream));
        dev_off += fill;
        fill = 0;
        pieces.clear();
        ++chunk;
        waited = false;
        return DFB_OK;
    };
    for (int i = 0; i < count; ++i) {
        size_t done = 0;
        while (done < bytes[i]) {
            if (!waited && chunk >= NG) DFB_CUDA(cudaEventSynchronize(ctx->gather_ev[chunk % NG]));
            waited = true;
            const size_t room = GS - fill;
            const size_t take = std::min(room, bytes[i] - done);
            char* base = pin + (chunk % NG) * GS + fill;
            for (size_t o = 0; o < take; o += (128u << 10))
                pieces.push_back({base + o, (const char*)srcs[i] + done + o, std::min<size_t>(128u << 10, take - o)});
            fill += take;
            done += take;
            if (fill == GS) DFB_TRY(flush());
        }
    }
    DFB_TRY(flush());
    // the sources have been copied into pinned memory: without `sync` the caller may go on (and may free
    // them) but must synchronise the stream before any other user of the ring runs
    if (sync) DFB_CUDA(cudaStreamSynchronize(ctx->stream));
    return DFB_OK;
}

// int32 -> uint16 with the conversion riding on the copy into pinned memory.  Returns true when a value does
// not fit (negative or > 65535).
static bool narrow16_stream(uint16_t* dst, const int32_t* src, size_t n) {
    uint32_t bad = 0;
    size_t i = 0;
    for (; i < n && ((uintptr_t)(dst + i) & 15); ++i) {
        bad |= (uint32_t)src[i] & 0xffff0000u;
        dst[i] = (uint16_t)src[i];
    }
    __m128i acc = _mm_setzero_si128();
    const __m128i bias = _mm_set1_epi32(32768), flip = _mm_set1_epi16((short)0x8000), hi = _mm_set1_epi32((int)0xffff0000u);
    for (; i + 16 <= n; i += 16) {
        const __m128i a = _mm_loadu_si128((const __m128i*)(src + i));
        const __m128i b = _mm_loadu_si128((const __m128i*)(src + i + 4));
        const __m128i c = _mm_loadu_si128((const __m128i*)(src + i + 8));
        const __m128i d = _mm_loadu_si128((const __m128i*)(src + i + 12));
        acc = _mm_or_si128(acc, _mm_and_si128(_mm_or_si128(_mm_or_si128(a, b), _mm_or_si128(c, d)), hi));
        // signed saturating pack of (x - 32768), then flip the sign bit: exact for 0 <= x <= 65535
        const __m128i p0 = _mm_xor_si128(_mm_packs_epi32(_mm_sub_epi32(a, bias), _mm_sub_epi32(b, bias)), flip);
        const __m128i p1 = _mm_xor_si128(_mm_packs_epi32(_mm_sub_epi32(c, bias), _mm_sub_epi32(d, bias)), flip);
        _mm_stream_si128((__m128i*)(dst + i), p0);
        _mm_stream_si128((__m128i*)(dst + i + 8), p1);
    }
    for (; i < n; ++i) {
        bad |= (uint32_t)src[i] & 0xffff0000u;
        dst[i] = (uint16_t)src[i];
    }
    _mm_sfence();
    alignas(16) uint32_t lanes[4];
    _mm_store_si128((__m128i*)lanes, acc);
    return (bad | lanes[0] | lanes[1] | lanes[2] | lanes[3]) != 0;
}

// h2d_gather for int32 columns that travel as uint16 (half the PCIe bytes): `counts[i]` elements of srcs[i] are
// narrowed into the pinned ring by the copy pool and DMAed to dst (uint16, columns back to back).  *overflow is
// set -- and the transfer abandoned -- as soon as a value does not fit; the caller then falls back to 32 bits.
int h2d_gather_narrow16(dfb_ctx* ctx, uint16_t* dst, const int32_t* const* srcs, const int64_t* counts, int count, bool sync,
                        bool* overflow) {
    *overflow = false;
    DFB_TRY(ensure_staging(ctx));
    constexpr size_t GS = SUB_BYTES, NG = NSUB;
    DFB_TRY(ensure_substage_events(ctx));
    char* pin = (char*)ctx->pinned;
    struct Piece { uint16_t* d; const int32_t* s; size_t n; };
    size_t& chunk = ctx->gather_chunk;
    size_t fill = 0, dev_off = 0;  // bytes
    std::vector<Piece> pieces;
    bool waited = false;
    std::atomic<int> bad{0};
    auto flush = [&]() -> int {
        if (fill == 0) return DFB_OK;
        ctx->pool->parallel_for((int)pieces.size(), [&](int i) {
            if (narrow16_stream(pieces[i].d, pieces[i].s, pieces[i].n)) bad.store(1, std::memory_order_relaxed);
        });
        DFB_CUDA(cudaMemcpyAsync((char*)dst + dev_off, pin + (chunk % NG) * GS, fill, cudaMemcpyHostToDevice, ctx->stream));
        DFB_CUDA(cudaEventRecord(ctx->gather_ev[chunk % NG], ctx->stream));
        dev_off += fill;
        fill = 0;
        pieces.clear();
        ++chunk;
        waited = false;
        return DFB_OK;
    };
    for (int i = 0; i < count && !bad.load(std::memory_order_relaxed); ++i) {
        size_t done = 0;  // elements
        const size_t total = (size_t)counts[i];
        while (done < total) {
            if (!waited && chunk >= NG) DFB_CUDA(cudaEventSynchronize(ctx->gather_ev[chunk % NG]));
            waited = true;
            const size_t room = (GS - fill) / 2;
            const size_t take = std::min(room, total - done);
            uint16_t* base = reinterpret_cast<uint16_t*>(pin + (chunk % NG) * GS + fill);
            for (size_t o = 0; o < take; o += (64u << 10))
                pieces.push_back({base + o, srcs[i] + done + o, std::min<size_t>(64u << 10, take - o)});
            fill += take * 2;
            done += take;
            if (fill + 2 > GS) DFB_TRY(flush());
        }
    }
    DFB_TRY(flush());
    *overflow = bad.load() != 0;
    if (sync || *overflow) DFB_CUDA(cudaStreamSynchronize(ctx->stream));
    return DFB_OK;
}

int h2d_big(dfb_ctx* ctx, void* dst, const void* src, size_t bytes) {
    if (bytes < (1u << 20) || is_pinned_host(src)) {
        DFB_TRY(h2d(ctx, dst, src, bytes));
        DFB_CUDA(cudaStreamSynchronize(ctx->stream));
        return DFB_OK;
    }
    DFB_TRY(ensure_staging(ctx));
    char* pin = (char*)ctx->pinned;
    const size_t nchunk = (bytes + STAGE_BYTES - 1) / STAGE_BYTES;
    for (size_t i = 0; i < nchunk; ++i) {
        const size_t off = i * STAGE_BYTES, sz = std::min(STAGE_BYTES, bytes - off);
        if (i >= NSTAGE) DFB_CUDA(cudaEventSynchronize(ctx->stage_ev[i % NSTAGE]));  // staging buffer free again
        pool_memcpy(ctx, pin + (i % NSTAGE) * STAGE_BYTES, (const char*)src + off, sz);
        DFB_CUDA(cudaMemcpyAsync((char*)dst + off, pin + (i % NSTAGE) * STAGE_BYTES, sz, cudaMemcpyHostToDevice,
                                 ctx->stream));
        DFB_CUDA(cudaEventRecord(ctx->stage_ev[i % NSTAGE], ctx->stream));
    }
    DFB_CUDA(cudaStreamSynchronize(ctx->stream));
    return DFB_OK;
}

// ------------------------------------------------------------------------------------------
// Exclusive scan, uint32 counts -> int64 offsets.  Three-phase (tile reduce, recursive scan of
// the tile sums, tile scan + offset).  Tiles of 2048 items (256 threads x 8), loads coalesced.

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

template <typename TIn>
__global__ void __launch_bounds__(SCAN_THREADS) scan_reduce_kernel(const TIn* __restrict__ in, int64_t n,
                                                                   int64_t* __restrict__ tile_sums) {
    __shared__ int64_t wsum[SCAN_THREADS / 32];
    int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
    int64_t acc = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        int64_t idx = base + (int64_t)i * SCAN_THREADS + threadIdx.x;
        if (idx < n) acc += (int64_t)in[idx];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        int64_t t = 0;
#pragma unroll
        for (int w = 0; w < SCAN_THREADS / 32; ++w) t += wsum[w];
        tile_sums[blockIdx.x] = t;
    }
}

// in-place exclusive scan of up to SCAN_TILE int64 values by one block; writes the total.
__global__ void __launch_bounds__(SCAN_THREADS) scan_small_kernel(int64_t* __restrict__ data, int64_t n,
                                                                  int64_t* __restrict__ total) {
    __shared__ int64_t wsum[SCAN_THREADS / 32];
    __shared__ int64_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    // each thread owns SCAN_ITEMS consecutive values (blocked arrangement)
    for (int64_t base = 0; base < n; base += SCAN_TILE) {
        int64_t v[SCAN_ITEMS];
        int64_t tsum = 0;
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; ++i) {
            int64_t idx = base + (int64_t)threadIdx.x * SCAN_ITEMS + i;
            v[i] = idx < n ? data[idx] : 0;
            tsum += v[i];
        }
        int64_t inc = tsum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int64_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if ((threadIdx.x & 31) >= o) inc += t;
        }
        if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = inc;
        __syncthreads();
        int64_t wprefix = 0;
        for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) wprefix += wsum[w];
        int64_t excl = carry_s + wprefix + inc - tsum;
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; ++i) {
            int64_t idx = base + (int64_t)threadIdx.x * SCAN_ITEMS + i;
            if (idx < n) data[idx] = excl;
            excl += v[i];
        }
        __syncthreads();
        if (threadIdx.x == SCAN_THREADS - 1) carry_s = excl;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total) *total = carry_s;
}

template <typename TIn>
__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_kernel(const TIn* __restrict__ in, int64_t n,
                                                                 const int64_t* __restrict__ tile_offsets,
                                                                 int64_t* __restrict__ out) {
    __shared__ int64_t wsum[SCAN_THREADS / 32];
    int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
    int64_t v[SCAN_ITEMS];
    int64_t tsum = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        int64_t idx = base + (int64_t)threadIdx.x * SCAN_ITEMS + i;
        v[i] = idx < n ? (int64_t)in[idx] : 0;
        tsum += v[i];
    }
    int64_t inc = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int64_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if ((threadIdx.x & 31) >= o) inc += t;
    }
    if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = inc;
    __syncthreads();
    int64_t wprefix = 0;
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) wprefix += wsum[w];
    int64_t excl = tile_offsets[blockIdx.x] + wprefix + inc - tsum;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        int64_t idx = base + (int64_t)threadIdx.x * SCAN_ITEMS + i;
        if (idx < n) out[idx] = excl;
        excl += v[i];
    }
}

// Whole scan by ONE block (small inputs: region tables, tile counts of a short chromosome): one launch and no
// scratch instead of reduce + scan-of-sums + tile scan.
template <typename TIn>
__global__ void __launch_bounds__(SCAN_THREADS) scan_one_block_kernel(const TIn* __restrict__ in, int64_t n,
                                                                      int64_t* __restrict__ out, int64_t* __restrict__ total) {
    __shared__ int64_t wsum[SCAN_THREADS / 32];
    __shared__ int64_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int64_t base = 0; base < n; base += SCAN_TILE) {
        int64_t v[SCAN_ITEMS];
        int64_t tsum = 0;
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; ++i) {
            const int64_t idx = base + (int64_t)threadIdx.x * SCAN_ITEMS + i;
            v[i] = idx < n ? (int64_t)in[idx] : 0;
            tsum += v[i];
        }
        int64_t inc = tsum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int64_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if ((threadIdx.x & 31) >= o) inc += t;
        }
        if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = inc;
        __syncthreads();
        int64_t wprefix = 0;
        for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) wprefix += wsum[w];
        int64_t excl = carry_s + wprefix + inc - tsum;
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; ++i) {
            const int64_t idx = base + (int64_t)threadIdx.x * SCAN_ITEMS + i;
            if (idx < n) out[idx] = excl;
            excl += v[i];
        }
        __syncthreads();
        if (threadIdx.x == SCAN_THREADS - 1) carry_s = excl;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total) *total = carry_s;
}

template <typename TIn>
static int scan_impl(dfb_ctx* ctx, const TIn* in, int64_t* out, int64_t n, int64_t* total_dev, int fam) {
    if (n <= 0) {
        if (total_dev) DFB_CUDA(cudaMemsetAsync(total_dev, 0, sizeof(int64_t), ctx->stream));
        return DFB_OK;
    }
    if (n <= (int64_t)SCAN_TILE * 16 && (const void*)in != (const void*)out) {
        KTimer t(ctx, fam, 1);
        scan_one_block_kernel<TIn><<<1, SCAN_THREADS, 0, ctx->stream>>>(in, n, out, total_dev);
        DFB_CUDA(cudaGetLastError());
        return DFB_OK;
    }
    int64_t tiles = ceil_div(n, SCAN_TILE);
    DevBuf<int64_t> sums;
    DFB_TRY(sums.alloc(ctx, (size_t)tiles));
    {
        KTimer t(ctx, fam, 1);
        scan_reduce_kernel<TIn><<<(unsigned)tiles, SCAN_THREADS, 0, ctx->stream>>>(in, n, sums.p);
    }
    if (tiles <= SCAN_TILE * 64) {
        KTimer t(ctx, fam, 1);
        scan_small_kernel<<<1, SCAN_THREADS, 0, ctx->stream>>>(sums.p, tiles, total_dev);
    } else {
        DevBuf<int64_t> sums_out;
        DFB_TRY(sums_out.alloc(ctx, (size_t)tiles));
        DFB_TRY(scan_impl<int64_t>(ctx, sums.p, sums_out.p, tiles, total_dev, fam));
        sums = std::move(sums_out);
    }
    {
        KTimer t(ctx, fam, 1);
        scan_tile_kernel<TIn><<<(unsigned)tiles, SCAN_THREADS, 0, ctx->stream>>>(in, n, sums.p, out);
    }
    DFB_CUDA(cudaGetLastError());
    return DFB_OK;
}

int exclusive_scan_u32(dfb_ctx* ctx, const uint32_t* in, int64_t* out, int64_t n, int64_t* total_dev, int fam) {
    return scan_impl<uint32_t>(ctx, in, out, n, total_dev, fam);
}

int exclusive_scan_i64(dfb_ctx* ctx, const int64_t* in, int64_t* out, int64_t n, int64_t* total_dev, int fam) {
    return scan_impl<int64_t>(ctx, in, out, n, total_dev, fam);
}

}  // namespace dfb

using namespace dfb;

extern "C" {

const char* dfb_last_error(void) { return g_err.c_str(); }
int dfb_version(void) { return DFB_VERSION; }

int dfb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int dfb_create(int device, void* stream, dfb_ctx** out) {
    DFB_REQUIRE(out != nullptr, "dfb_create: out is NULL");
    *out = nullptr;
    int ndev = dfb_device_count();
    if (ndev <= 0) {
        set_error("dfb_create: no CUDA device available (this engine has no CPU fallback)");
        return DFB_ECUDA;
    }
    DFB_REQUIRE(device >= 0 && device < ndev, "dfb_create: device %d out of range (%d visible)", device, ndev);
    DFB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    DFB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        set_error("dfb_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only", device,
                  prop.major, prop.minor);
        return DFB_ECUDA;
    }
    dfb_ctx* c = new dfb_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->smem_optin = prop.sharedMemPerBlockOptin;
    if (stream) {
        c->stream = (cudaStream_t)stream;
    } else {
        cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) {
            set_error("cudaStreamCreate: %s", cudaGetErrorString(e));
            delete c;
            return DFB_ECUDA;
        }
        c->own_stream = true;
    }
    // A PRIVATE stream-ordered pool that keeps freed blocks (repeated calls re-use them without a device
    // synchronisation).  The device's default pool is left alone: other cudaMallocAsync users of the host
    // process keep their own release policy.
    {
        cudaMemPoolProps props;
        memset(&props, 0, sizeof props);
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = device;
        if (cudaMemPoolCreate(&c->mempool, &props) == cudaSuccess) {
            uint64_t thr = UINT64_MAX;
            cudaMemPoolSetAttribute(c->mempool, cudaMemPoolAttrReleaseThreshold, &thr);
        } else {
            c->mempool = nullptr;  // fall back to the default pool with its default policy
        }
    }
    cudaGetLastError();
    if (const char* e = getenv("DFB_POISON")) c->opt_poison = atoi(e) != 0;
    *out = c;
    return DFB_OK;
}

void dfb_destroy(dfb_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    dfb_comm_destroy(c);
    for (auto& e : c->evs) {
        cudaEventDestroy(e.a);
        cudaEventDestroy(e.b);
    }
    bg_download_wait(c);
    if (c->bg.stream) {
        cudaStreamSynchronize(c->bg.stream);
        cudaStreamDestroy(c->bg.stream);
        cudaEventDestroy(c->bg.ready);
        cudaEventDestroy(c->bg.ev[0]);
        cudaEventDestroy(c->bg.ev[1]);
        cudaFreeHost(c->bg.pin);
    }
    for (auto e : c->ev_free) cudaEventDestroy(e);
    for (auto e : c->gather_ev)
        if (e) cudaEventDestroy(e);
    if (c->pinned) {
        cudaFreeHost(c->pinned);
        for (int i = 0; i < 4; ++i) cudaEventDestroy(c->stage_ev[i]);
    }
    if (c->aux) {
        cudaStreamSynchronize(c->aux);
        cudaStreamDestroy(c->aux);
        cudaEventDestroy(c->ev_main);
        cudaEventDestroy(c->ev_aux);
    }
    for (auto& b : c->pin_blocks) cudaFreeHost(b.p);
    if (c->lut_dev) cudaFree(c->lut_dev);
    if (c->k2_mask) cudaFree(c->k2_mask);
    if (c->k2_gflag) cudaFree(c->k2_gflag);
    if (c->mempool) cudaMemPoolDestroy(c->mempool);  // released once the last outstanding block is freed
    delete c->pool;
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
}

int dfb_sync(dfb_ctx* c) {
    DFB_REQUIRE(c, "dfb_sync: NULL context");
    DFB_TRY(bg_download_wait(c));
    DFB_CUDA(cudaStreamSynchronize(c->stream));
    return DFB_OK;
}

int64_t dfb_launch_count(dfb_ctx* c) { return c ? c->launches : 0; }

static int drain_events(dfb_ctx* c) {
    if (c->evs.empty()) return DFB_OK;
    DFB_CUDA(cudaStreamSynchronize(c->stream));
    if (getenv("DFB_GAPS")) {
        // developer aid: device idle time between consecutive timed launches (host-induced bubbles)
        double busy = 0, idle = 0;
        for (size_t i = 0; i < c->evs.size(); ++i) {
            float ms = 0.f, gap = 0.f;
            cudaEventElapsedTime(&ms, c->evs[i].a, c->evs[i].b);
            busy += ms;
            if (i + 1 < c->evs.size()) cudaEventElapsedTime(&gap, c->evs[i].b, c->evs[i + 1].a);
            idle += gap;
            const char* f = strrchr(c->evs[i].file, '/');
            fprintf(stderr, "[dfb] %-14s:%-5d fam %d  run %8.3f ms   then idle %8.3f ms\n", f ? f + 1 : c->evs[i].file,
                    c->evs[i].line, c->evs[i].k, ms, gap);
        }
        fprintf(stderr, "[dfb] busy %.3f ms, idle between timed launches %.3f ms\n", busy, idle);
    }
    for (auto& e : c->evs) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, e.a, e.b) == cudaSuccess) c->kms[e.k] += ms;
        c->ev_free.push_back(e.a);
        c->ev_free.push_back(e.b);
    }
    c->evs.clear();
    cudaGetLastError();
    return DFB_OK;
}

void dfb_reset_counters(dfb_ctx* c) {
    if (!c) return;
    drain_events(c);
    c->launches = 0;
    for (int k = 0; k < DFB_K_COUNT; ++k) {
        c->kms[k] = 0;
        c->klaunch[k] = 0;
    }
}

int dfb_enable_timing(dfb_ctx* c, int on) {
    DFB_REQUIRE(c, "NULL context");
    DFB_TRY(drain_events(c));
    c->timing = on == 0 ? 0u : (on == 1 ? ~0u : (unsigned)on);  // 1 = every family, else a mask of (1 << family)
    return DFB_OK;
}

int dfb_kernel_time(dfb_ctx* c, int kernel, double* ms, int64_t* launches) {
    DFB_REQUIRE(c && kernel > 0 && kernel < DFB_K_COUNT, "dfb_kernel_time: bad kernel id %d", kernel);
    DFB_TRY(drain_events(c));
    if (ms) *ms = c->kms[kernel];
    if (launches) *launches = c->klaunch[kernel];
    return DFB_OK;
}

int dfb_set_option(dfb_ctx* c, const char* key, double value) {
    DFB_REQUIRE(c && key, "dfb_set_option: NULL argument");
    if (!strcmp(key, "screen")) c->opt_screen = value != 0;
    else if (!strcmp(key, "fused")) c->opt_fused = value != 0;
    else if (!strcmp(key, "null2")) c->opt_null2 = (int)value;  // 0 off, 1 auto (more than 16 samples), 2 always
    else if (!strcmp(key, "mma_warps")) c->opt_mma_warps = (int)value;
    else if (!strcmp(key, "a_stages")) c->opt_a_stages = (int)value;
    else if (!strcmp(key, "a16_fused")) c->opt_a16_fused = value != 0;
    else if (!strcmp(key, "lazy_y")) c->opt_lazy_y = value != 0;
    else if (!strcmp(key, "mask_cache")) c->opt_mask_cache = value != 0;
    else if (!strcmp(key, "filter_events")) c->opt_filter_events = value != 0;
    else if (!strcmp(key, "wire16")) c->opt_wire16 = value != 0;
    else if (!strcmp(key, "pinned_direct")) c->opt_pinned_direct = value != 0;
    else if (!strcmp(key, "cluster_one_block")) c->opt_cluster_one_block = value != 0;
    else if (!strcmp(key, "dbg")) c->opt_dbg = (int)value;
    else if (!strcmp(key, "tmem_stages")) c->opt_tmem_stages = (int)value;
    else if (!strcmp(key, "k2_waves")) c->opt_k2_waves = (int)value;
    else if (!strcmp(key, "chunk_perms")) c->opt_chunk_perms = (int)value;
    else if (!strcmp(key, "ys_smem")) c->opt_ys_smem = (int)value;
    else if (!strcmp(key, "b_stages")) c->opt_b_stages = (int)value;
    else if (!strcmp(key, "epi_warps")) c->opt_epi_warps = (int)value;
    else if (!strcmp(key, "null_capacity_mb")) c->opt_null_cap_mb = value;
    else if (!strcmp(key, "perm_batch")) c->opt_perm_batch = (int)value;
    else if (!strcmp(key, "scratch_gb")) c->opt_scratch_gb = value;
    else {
        set_error("dfb_set_option: unknown option '%s'", key);
        return DFB_EINVAL;
    }
    return DFB_OK;
}

}  // extern "C"
