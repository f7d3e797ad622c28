// K3 / K4 -- threshold, run segmentation, per-region area, genomic mapping and maxGap clustering.
//
// Reference: findRegions (R/findRegions.R:116-296), .getSegmentsRle (:363-396, IRanges::slice with
// inclusive bounds), .clusterMakerRle (:445-475), and the Views sum/mean of IRanges over a numeric
// Rle (run-by-run accumulation; see view_sums in oracle/derfinder_oracle.py).
//
// Everything works on bit masks: one bit per kept base and row (row = observed statistic, or one
// permutation [x side]).  A region is a maximal run of set bits that does not cross a set bit of
// the segment-start mask SEG (segments = .clusterMakerRle(position, maxRegionGap, ranges=TRUE),
// R/calculatePvalues.R:234-237).  Region starts are found with word-level bit tricks, counted with
// popc, placed with a device-wide exclusive scan (= (row, index) order, the reference's order) and
// then every region is summed left to right by one thread in IRanges' order: a run of k equal
// adjacent values contributes value*k as ONE product (bit-exact areas).
#include <math.h>

#include <algorithm>

#include "data.cuh"

namespace dfb {

}  // namespace dfb

// ------------------------------------------------------------------------------------------
// position index

using dfb::h2d;
using dfb::set_error;

int PositionIndex::build(dfb_ctx* ctx, const int32_t* pos_len, int64_t nruns, int pos_first) {
    // This runs on the host at the start of every chromosome while the device has little queued, so
    // the common case -- runs already normalised (all lengths > 0, values alternating) -- is a plain
    // copy plus one pass over the TRUE runs into pre-sized arrays.
    len.clear();
    first = pos_first ? 1 : 0;
    chr_len = 0;
    int32_t lo = 1;
    int64_t total = 0;
    for (int64_t r = 0; r < nruns; ++r) {
        lo = std::min(lo, pos_len[r]);
        total += pos_len[r];
    }
    DFB_REQUIRE(lo >= 0, "position: negative run length");
    if (lo > 0) {
        len.assign(pos_len, pos_len + nruns);
        chr_len = total;
    } else {
        // normalise: drop zero-length runs, merge neighbours of equal value
        int cur_val = first;
        bool have = false;
        int last_val = 0;
        for (int64_t r = 0; r < nruns; ++r, cur_val ^= 1) {
            if (pos_len[r] == 0) continue;
            if (have && last_val == cur_val) {
                DFB_REQUIRE((int64_t)len.back() + pos_len[r] <= INT32_MAX, "position run longer than 2^31-1");
                len.back() += pos_len[r];
            } else {
                if (!have) first = cur_val;
                len.push_back(pos_len[r]);
                last_val = cur_val;
                have = true;
            }
            chr_len += pos_len[r];
        }
    }
    DFB_REQUIRE(chr_len <= INT32_MAX, "chromosome length %lld exceeds R's integer range", (long long)chr_len);
    const size_t nl = len.size();
    const size_t r0 = first ? 0 : 1;  // index of the first TRUE run
    K = nl > r0 ? (int64_t)((nl - r0 + 1) / 2) : 0;
    h_gstart.resize((size_t)K);
    h_cum.resize((size_t)K + 1);
    int64_t gpos = 1 + ((r0 == 1 && nl > 0) ? (int64_t)len[0] : 0), kept = 0;
    const int32_t* lp = len.data();
    for (size_t k = 0, r = r0; r < nl; ++k, r += 2) {
        h_gstart[k] = (int32_t)gpos;
        h_cum[k] = kept;
        kept += lp[r];
        gpos += (int64_t)lp[r] + (r + 1 < nl ? (int64_t)lp[r + 1] : 0);
    }
    h_cum[(size_t)K] = kept;
    N = kept;
    DFB_TRY(run_gstart.alloc(ctx, (size_t)std::max<int64_t>(K, 1)));
    DFB_TRY(run_cum.alloc(ctx, (size_t)K + 1));
    // one pinned staging block for both tables: a pageable cudaMemcpyAsync first drains the stream and
    // then goes through the driver's bounce buffer, twice
    const size_t b_g = (size_t)K * sizeof(int32_t), b_c = (size_t)(K + 1) * sizeof(int64_t);
    unsigned char* pin = (unsigned char*)dfb::arena_alloc(ctx, b_c + b_g);
    if (pin) {
        memcpy(pin, h_cum.data(), b_c);
        memcpy(pin + b_c, h_gstart.data(), b_g);
        DFB_TRY(h2d(ctx, run_cum.p, pin, b_c));
        DFB_TRY(h2d(ctx, run_gstart.p, pin + b_c, b_g));
    } else {
        DFB_TRY(h2d(ctx, run_gstart.p, h_gstart.data(), b_g));
        DFB_TRY(h2d(ctx, run_cum.p, h_cum.data(), b_c));
    }
    const cudaError_t e = cudaStreamSynchronize(ctx->stream);  // host vectors may be reallocated by the caller
    if (pin) dfb::arena_release(ctx, pin, b_c + b_g);
    DFB_CUDA(e);
    return DFB_OK;
}

namespace dfb {

// ------------------------------------------------------------------------------------------
// segment starts: TRUE run k opens a new segment iff it is the first run or the genomic gap to
// the previous TRUE run exceeds maxRegionGap (reduce(ir, min.gapwidth = maxGap + 1), :454).

__global__ void segment_start_kernel(const int32_t* __restrict__ gstart, const int64_t* __restrict__ cum,
                                     int64_t K, int32_t max_gap, uint32_t* __restrict__ seg) {
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    bool open = true;
    if (k > 0) {
        int64_t prev_end = (int64_t)gstart[k - 1] + (cum[k] - cum[k - 1]) - 1;
        int64_t gap = (int64_t)gstart[k] - prev_end - 1;
        open = gap > (int64_t)max_gap;
    }
    if (open) {
        int64_t idx = cum[k];
        atomicOr(&seg[idx >> 5], 1u << (idx & 31));
    }
}

int build_segment_starts(dfb_ctx* ctx, const PositionIndex& pos, int32_t max_region_gap, DevBuf<uint32_t>* seg) {
    int64_t W = ceil_div(pos.N, 32);
    DFB_TRY(seg->alloc(ctx, (size_t)std::max<int64_t>(W, 1)));
    DFB_TRY(seg->zero());
    if (pos.K > 0) {
        KTimer t(ctx, DFB_K_REGIONS);
        segment_start_kernel<<<(unsigned)ceil_div(pos.K, 256), 256, 0, ctx->stream>>>(
            pos.run_gstart.p, pos.run_cum.p, pos.K, max_region_gap, seg->p);
    }
    DFB_CUDA(cudaGetLastError());
    return DFB_OK;
}

// ------------------------------------------------------------------------------------------
// threshold mask of a dense vector: row 0 = x >= U ("up"), row 1 = x <= L ("dn") -- inclusive,
// IRanges::slice defaults (R/findRegions.R:384-391).  NaN never passes.

__global__ void __launch_bounds__(256) threshold_mask_kernel(const double* __restrict__ x, int64_t N, double U,
                                                             double L, uint32_t* __restrict__ up,
                                                             uint32_t* __restrict__ dn, int64_t W) {
    const int lane = threadIdx.x & 31;
    int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t w = warp; w < W; w += nwarps) {
        const int64_t i = w * 32 + lane;
        const double v = i < N ? x[i] : nan("");
        const unsigned mu = __ballot_sync(0xffffffffu, v >= U);
        const unsigned md = __ballot_sync(0xffffffffu, v <= L);
        if (lane == 0) {
            up[w] = mu;
            if (dn) dn[w] = md;
        }
    }
}

// ------------------------------------------------------------------------------------------
// region starts per word

__device__ __forceinline__ uint32_t region_starts(uint32_t m, uint32_t prev_msb, uint32_t seg) {
    return m & (~((m << 1) | prev_msb) | seg);
}

struct EmitParams {
    const uint32_t* mask;
    const uint32_t* seg;
    int64_t rows, W, N;
    int sides;  // rows per permutation (1 or 2)
    // value source
    const double* dense;  // [rows_dense][N] (row stride N) when not compact
    int64_t dense_stride;
    const int64_t* off;  // compact: [rows][tiles]
    int64_t tiles;
    int tile_bases;
    const double* vals;
    // outputs (any may be null)
    double* area;
    double* stat;
    int32_t* width;
    int32_t* istart;
    int32_t* iend;
    int32_t* perm;  // 1-based permutation id = row / sides + 1
};

// position (global word index << 5 | bit) of every region start, in (row, index) order:
// The null scan has rows * W ~ 10^7 mask words but ~10^6 region starts: counting per GROUP of 32 words
// (one warp, row-local so that a row starts a group) keeps the scan 32x smaller and drops the per-word
// count and offset arrays (180 B of traffic per 32 words -> 12); the fill pass recomputes the word's
// starts and places them with a warp prefix sum.
constexpr int GROUPS_PER_WARP = 8;  // independent 128-byte loads a warp keeps in flight

// `gflag` (optional, [rows][Gr]): 0 marks a group without a single passing base -- at stringent cutoffs
// nearly all of them -- whose mask words are then not even loaded.
__global__ void __launch_bounds__(256) count_starts_group_kernel(const uint32_t* __restrict__ mask,
                                                                 const uint32_t* __restrict__ seg, int64_t rows,
                                                                 int64_t W, int64_t Gr, uint32_t* __restrict__ counts,
                                                                 const unsigned char* __restrict__ gflag) {
    // grid: (groups of a row / 64, rows); a warp owns GROUPS_PER_WARP consecutive groups and issues all
    // of their loads before it looks at any (one load per short-lived warp left the kernel latency-bound)
    const int lane = threadIdx.x & 31;
    for (int64_t row = blockIdx.y; row < rows; row += gridDim.y) {
        const int64_t k0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * GROUPS_PER_WARP;
        if (k0 >= Gr) continue;
        unsigned live = 0xffu;  // bit u: group k0 + u may hold passing bases
        if (gflag) {
            const bool f = lane < GROUPS_PER_WARP && k0 + lane < Gr && gflag[row * Gr + k0 + lane] != 0;
            live = __ballot_sync(0xffffffffu, f);
            if (live == 0u) {
                if (lane < GROUPS_PER_WARP && k0 + lane < Gr) counts[row * Gr + k0 + lane] = 0u;
                continue;
            }
        }
        uint32_t m[GROUPS_PER_WARP], pv[GROUPS_PER_WARP], sg[GROUPS_PER_WARP];
#pragma unroll
        for (int u = 0; u < GROUPS_PER_WARP; ++u) {
            const int64_t w = (k0 + u) * 32 + lane;
            const bool valid = w < W && ((live >> u) & 1u);
            const int64_t g = row * W + w;
            m[u] = valid ? mask[g] : 0u;
            pv[u] = (valid && w > 0) ? mask[g - 1] : 0u;
            sg[u] = valid ? seg[w] : 0u;
        }
#pragma unroll
        for (int u = 0; u < GROUPS_PER_WARP; ++u) {
            uint32_t c = __popc(region_starts(m[u], pv[u] >> 31, sg[u]));
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
            if (lane == 0 && k0 + u < Gr) counts[row * Gr + k0 + u] = c;
        }
    }
}

__global__ void __launch_bounds__(256) fill_starts_group_kernel(const uint32_t* __restrict__ mask,
                                                                const uint32_t* __restrict__ seg,
                                                                const int64_t* __restrict__ goffsets, int64_t rows,
                                                                int64_t W, int64_t Gr, int64_t* __restrict__ start_pos,
                                                                const unsigned char* __restrict__ gflag) {
    const int lane = threadIdx.x & 31;
    for (int64_t row = blockIdx.y; row < rows; row += gridDim.y) {
        const int64_t k0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * GROUPS_PER_WARP;
        if (k0 >= Gr) continue;
        unsigned live = 0xffu;
        if (gflag) {
            const bool f = lane < GROUPS_PER_WARP && k0 + lane < Gr && gflag[row * Gr + k0 + lane] != 0;
            live = __ballot_sync(0xffffffffu, f);
            if (live == 0u) continue;
        }
        uint32_t m[GROUPS_PER_WARP], pv[GROUPS_PER_WARP], sg[GROUPS_PER_WARP];
#pragma unroll
        for (int u = 0; u < GROUPS_PER_WARP; ++u) {
            const int64_t w = (k0 + u) * 32 + lane;
            const bool valid = w < W && ((live >> u) & 1u);
            const int64_t g = row * W + w;
            m[u] = valid ? mask[g] : 0u;
            pv[u] = (valid && w > 0) ? mask[g - 1] : 0u;
            sg[u] = valid ? seg[w] : 0u;
        }
#pragma unroll
        for (int u = 0; u < GROUPS_PER_WARP; ++u) {
            uint32_t starts = region_starts(m[u], pv[u] >> 31, sg[u]);
            if (!__any_sync(0xffffffffu, starts != 0u)) continue;
            const int c = __popc(starts);
            int inc = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += t;
            }
            const int64_t g = row * W + (k0 + u) * 32 + lane;
            int64_t out = goffsets[row * Gr + k0 + u] + (int64_t)(inc - c);
            while (starts) {
                const int b = __ffs(starts) - 1;
                starts &= starts - 1;
                start_pos[out++] = (g << 5) | (int64_t)b;
            }
        }
    }
}

// Flag-driven variants for the tensor-core null scan, which leaves a byte per group (1 = some word of the group
// is non-empty).  At stringent cutoffs nearly every group is empty: a warp reads the flags of 128 consecutive
// groups per step (four per lane), zeroes the counts of the empty ones and visits the live ones one by one --
// the word-parallel kernels above spent one warp per eight groups just to find out that there was nothing to do.
constexpr int FLAG_GROUPS_PER_WARP = 128;

__global__ void __launch_bounds__(256) count_starts_flagged_kernel(const uint32_t* __restrict__ mask,
                                                                   const uint32_t* __restrict__ seg, int64_t rows,
                                                                   int64_t W, int64_t Gr, uint32_t* __restrict__ counts,
                                                                   const unsigned char* __restrict__ gflag) {
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (int64_t row = blockIdx.y; row < rows; row += gridDim.y) {
        for (int64_t k0 = ((int64_t)blockIdx.x * wpb + (threadIdx.x >> 5)) * FLAG_GROUPS_PER_WARP; k0 < Gr;
             k0 += (int64_t)gridDim.x * wpb * FLAG_GROUPS_PER_WARP) {
            unsigned f = 0u;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int64_t k = k0 + 4 * lane + j;
                if (k < Gr) {
                    if (gflag[row * Gr + k]) f |= 1u << j;
                    else counts[row * Gr + k] = 0u;
                }
            }
            unsigned any = __ballot_sync(0xffffffffu, f != 0u);
            while (any) {
                const int src = __ffs(any) - 1;
                any &= any - 1;
                unsigned fs = __shfl_sync(0xffffffffu, f, src);
                while (fs) {
                    const int j = __ffs(fs) - 1;
                    fs &= fs - 1;
                    const int64_t k = k0 + 4 * src + j;
                    const int64_t w = k * 32 + lane;
                    const bool valid = w < W;
                    const int64_t g = row * W + w;
                    const uint32_t m = valid ? mask[g] : 0u;
                    const uint32_t pv = (valid && w > 0) ? mask[g - 1] : 0u;
                    const uint32_t sg = valid ? seg[w] : 0u;
                    uint32_t c = __popc(region_starts(m, pv >> 31, sg));
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
                    if (lane == 0) counts[row * Gr + k] = c;
                }
            }
        }
    }
}

__global__ void __launch_bounds__(256) fill_starts_flagged_kernel(const uint32_t* __restrict__ mask,
                                                                  const uint32_t* __restrict__ seg,
                                                                  const int64_t* __restrict__ goffsets, int64_t rows,
                                                                  int64_t W, int64_t Gr, int64_t* __restrict__ start_pos,
                                                                  const unsigned char* __restrict__ gflag) {
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (int64_t row = blockIdx.y; row < rows; row += gridDim.y) {
        for (int64_t k0 = ((int64_t)blockIdx.x * wpb + (threadIdx.x >> 5)) * FLAG_GROUPS_PER_WARP; k0 < Gr;
             k0 += (int64_t)gridDim.x * wpb * FLAG_GROUPS_PER_WARP) {
            unsigned f = 0u;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int64_t k = k0 + 4 * lane + j;
                if (k < Gr && gflag[row * Gr + k]) f |= 1u << j;
            }
            unsigned any = __ballot_sync(0xffffffffu, f != 0u);
            while (any) {
                const int src = __ffs(any) - 1;
                any &= any - 1;
                unsigned fs = __shfl_sync(0xffffffffu, f, src);
                while (fs) {
                    const int j = __ffs(fs) - 1;
                    fs &= fs - 1;
                    const int64_t k = k0 + 4 * src + j;
                    const int64_t w = k * 32 + lane;
                    const bool valid = w < W;
                    const int64_t g = row * W + w;
                    const uint32_t m = valid ? mask[g] : 0u;
                    const uint32_t pv = (valid && w > 0) ? mask[g - 1] : 0u;
                    const uint32_t sg = valid ? seg[w] : 0u;
                    uint32_t starts = region_starts(m, pv >> 31, sg);
                    if (!__any_sync(0xffffffffu, starts != 0u)) continue;
                    const int c = __popc(starts);
                    int inc = c;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int t = __shfl_up_sync(0xffffffffu, inc, o);
                        if (lane >= o) inc += t;
                    }
                    int64_t out = goffsets[row * Gr + k] + (int64_t)(inc - c);
                    while (starts) {
                        const int b = __ffs(starts) - 1;
                        starts &= starts - 1;
                        start_pos[out++] = (g << 5) | (int64_t)b;
                    }
                }
            }
        }
    }
}

// One thread per region: all lanes of a warp walk regions of similar length that are adjacent in
// memory (the word-parallel predecessor left most lanes idle while one walked).
// A region that runs past LONG_WORDS mask words is handed to walk_long_regions_kernel (one WARP per
// region) through `long_list`: the kernel's duration used to be the serial walk of its longest regions
// (thousands of bases on one thread) rather than its memory traffic.
constexpr int LONG_WORDS = 8;
template <bool COMPACT>
__global__ void __launch_bounds__(128) walk_regions_kernel(const EmitParams P, const int64_t* __restrict__ start_pos,
                                                           int64_t total, int32_t* __restrict__ long_list,
                                                           unsigned int* __restrict__ long_count) {
    const int64_t out = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (out >= total) return;
    const int64_t sp = start_pos[out];
    const int64_t g = sp >> 5;
    const int b = (int)(sp & 31);
    const int64_t row = g / P.W, w = g - row * P.W;
    const uint32_t m = P.mask[g];
    const uint32_t sg = P.seg[w];
    const uint32_t* mrow = P.mask + row * P.W;
    const int wpt_shift = COMPACT ? (31 - __clz(P.tile_bases >> 5)) : 0;  // log2(words per value tile)
    const int64_t i = w * 32 + b;
    // first value of the region
    const double* src;
    if (COMPACT) {
        const int64_t t = w >> wpt_shift;
        int64_t ptr = P.off[row * P.tiles + t];
        for (int64_t ww = t << wpt_shift; ww < w; ++ww) ptr += __popc(mrow[ww]);
        ptr += __popc(m & ((1u << b) - 1u));
        src = P.vals + ptr;
    } else {
        src = P.dense + row * P.dense_stride + i;
    }
    // The region is walked word by word: inside a word its extent is the run of bits that pass and do
    // not open a new segment (bit tricks, no per-base tests); values of consecutive passing bases are
    // consecutive in memory until a value-tile boundary.  Sum in IRanges' order: a run of k equal
    // adjacent values contributes value*k as ONE product.
    double cur = src[0];
    double runlen = 1.0;
    double acc = 0.0;
    int64_t cw = w;
    int64_t j = i + 1;  // one past the last base taken so far
    uint32_t cont = (b < 31) ? ((m & ~sg) >> (b + 1)) : 0u;  // bits above b: passing, not a segment start
    int avail = 31 - b;                                      // bits of `cont` that are real
    const double* vp = src + 1;
    for (;;) {
        int take = __ffs(~cont) - 1;  // trailing ones (cont has zeros above `avail`)
        if (take < 0 || take > avail) take = avail;
        for (int t2 = 0; t2 < take; ++t2) {
            const double v = vp[t2];
            if (v == cur) {
                runlen += 1.0;
            } else {
                acc = __dadd_rn(acc, __dmul_rn(cur, runlen));
                cur = v;
                runlen = 1.0;
            }
        }
        j += take;
        vp += take;
        if (take < avail || j >= P.N) break;  // the run ended inside this word
        if (long_list && cw - w >= LONG_WORDS) {
            long_list[atomicAdd(long_count, 1u)] = (int32_t)out;
            return;
        }
        ++cw;
        const uint32_t nm = mrow[cw], ns = P.seg[cw];
        cont = nm & ~ns;
        avail = 32;
        if (COMPACT && (cw & ((1 << wpt_shift) - 1)) == 0) vp = P.vals + P.off[row * P.tiles + (cw >> wpt_shift)];
    }
    acc = __dadd_rn(acc, __dmul_rn(cur, runlen));
    const int32_t wd = (int32_t)(j - i);
    if (P.area) P.area[out] = fabs(acc);
    if (P.stat) P.stat[out] = __ddiv_rn(acc, (double)wd);
    if (P.width) P.width[out] = wd;
    if (P.istart) P.istart[out] = (int32_t)(i + 1);
    if (P.iend) P.iend[out] = (int32_t)j;
    if (P.perm) P.perm[out] = (int32_t)(row / P.sides) + 1;
}

// One warp per long region: 32 bases (one mask word) per step, loaded together; run heads by ballot,
// per-run value * length products in parallel, one ordered addition per run -- the same IRanges order
// as the thread walk (and as view_sums_kernel), with ~1-2 serial steps per word instead of 32.
template <bool COMPACT>
__global__ void __launch_bounds__(128) walk_long_regions_kernel(const EmitParams P,
                                                                const int64_t* __restrict__ start_pos,
                                                                const int32_t* __restrict__ long_list,
                                                                const unsigned int* __restrict__ long_count) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t n_long = (int64_t)*long_count;
    const int wpt_shift = COMPACT ? (31 - __clz(P.tile_bases >> 5)) : 0;
    for (int64_t li = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; li < n_long; li += warps) {
        const int64_t out = long_list[li];
        const int64_t sp = start_pos[out];
        const int64_t g0 = sp >> 5;
        const int b0 = (int)(sp & 31);
        const int64_t row = g0 / P.W, w0 = g0 - row * P.W;
        const uint32_t* mrow = P.mask + row * P.W;
        const int64_t i0 = w0 * 32 + b0;
        double acc = 0.0, open_v = 0.0;
        int64_t open_len = 0, width = 0;
        // first value slot of a word (COMPACT): tile offset + passing bases of the tile's earlier words
        auto word_ptr = [&](int64_t cw) -> int64_t {
            const int64_t t = cw >> wpt_shift;
            int64_t ptr = P.off[row * P.tiles + t];
            for (int64_t ww = t << wpt_shift; ww < cw; ++ww) ptr += __popc(mrow[ww]);
            return ptr;
        };
        // the mask word, segment word and value offset of the NEXT word are requested before the values of
        // the current one: one dependent memory round trip per 32 bases instead of three
        uint32_t m = mrow[w0], sgw = P.seg[w0];
        int64_t wptr = COMPACT ? word_ptr(w0) : 0;
        for (int64_t cw = w0;; ++cw) {
            const bool has_next = (cw + 1) * 32 < P.N;
            uint32_t m_next = 0u, sg_next = 0u;
            int64_t wptr_next = 0;
            if (has_next) {
                m_next = mrow[cw + 1];
                sg_next = P.seg[cw + 1];
                if (COMPACT) wptr_next = word_ptr(cw + 1);
            }
            const int sbit = cw == w0 ? b0 : 0;
            // bits of the region inside this word: from sbit, while passing and not opening a new segment
            // (the region's own first base may sit on a segment start)
            uint32_t x = m & ~sgw;
            if (cw == w0) x |= 1u << b0;
            x >>= sbit;
            const int avail = (int)min((int64_t)(32 - sbit), P.N - (cw * 32 + sbit));
            int cnt = __ffs(~x) - 1;  // trailing ones
            if (cnt < 0 || cnt > avail) cnt = avail;
            const int bit = sbit + lane;
            const bool valid = lane < cnt;
            double v = 0.0;
            if (valid) {
                if (COMPACT) {
                    v = P.vals[wptr + __popc(m & ((1u << bit) - 1u))];
                } else {
                    v = P.dense[row * P.dense_stride + cw * 32 + bit];
                }
            }
            double pv = __shfl_up_sync(full, v, 1);
            if (lane == 0) pv = open_v;
            const bool head = valid && ((cw == w0 && lane == 0) || v != pv);
            const unsigned hm = __ballot_sync(full, head);
            if (hm == 0) {
                open_len += cnt;
            } else {
                open_len += __ffs(hm) - 1;
                if (open_len > 0) acc = __dadd_rn(acc, __dmul_rn(open_v, (double)open_len));
                const unsigned above = hm & ~((2u << lane) - 1u);
                const int next_head = above ? __ffs(above) - 1 : cnt;
                const double prod = __dmul_rn(v, (double)(next_head - lane));
                const int last = 31 - __clz(hm);
                unsigned mm = hm & ~(1u << last);
                while (mm) {
                    const int b = __ffs(mm) - 1;
                    mm &= mm - 1;
                    acc = __dadd_rn(acc, __shfl_sync(full, prod, b));
                }
                open_v = __shfl_sync(full, v, last);
                open_len = cnt - last;
            }
            width += cnt;
            if (cnt < 32 - sbit || !has_next) break;  // ended inside this word / at the last base
            m = m_next;
            sgw = sg_next;
            wptr = wptr_next;
        }
        if (open_len > 0) acc = __dadd_rn(acc, __dmul_rn(open_v, (double)open_len));
        if (lane == 0) {
            const int32_t wd = (int32_t)width;
            if (P.area) P.area[out] = fabs(acc);
            if (P.stat) P.stat[out] = __ddiv_rn(acc, (double)wd);
            if (P.width) P.width[out] = wd;
            if (P.istart) P.istart[out] = (int32_t)(i0 + 1);
            if (P.iend) P.iend[out] = (int32_t)(i0 + width);
            if (P.perm) P.perm[out] = (int32_t)(row / P.sides) + 1;
        }
    }
}

int find_regions_rows(dfb_ctx* ctx, int64_t N, int64_t rows, int sides, const uint32_t* mask, const uint32_t* seg,
                      const double* dense, int64_t dense_stride, const PermBatch* compact, bool want_index,
                      bool want_perm, RegionSet* out) {
    const int64_t W = ceil_div(N, 32);
    const int64_t total_words = rows * W;
    out->total = 0;
    out->row_counts.assign((size_t)rows, 0);
    if (total_words == 0) return DFB_OK;
    const int64_t Gr = ceil_div(W, 32), groups = rows * Gr;  // groups of 32 words, row-local
    const unsigned char* gflag = (compact && compact->gflag.p) ? compact->gflag.p : nullptr;
    const dim3 group_grid((unsigned)ceil_div(Gr, 8 * GROUPS_PER_WARP), (unsigned)std::min<int64_t>(rows, 65535));
    // flag-driven kernels: a block covers 8 * 128 groups of a row per step; enough blocks per row to fill the device
    const dim3 flag_grid((unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(Gr, 8 * FLAG_GROUPS_PER_WARP),
                                                                         ceil_div((int64_t)ctx->sm_count * 8, rows))),
                         (unsigned)std::min<int64_t>(rows, 65535));
    DevBuf<uint32_t> counts;
    DevBuf<int64_t> offsets, total_dev;
    DFB_TRY(counts.alloc(ctx, (size_t)groups));
    DFB_TRY(offsets.alloc(ctx, (size_t)groups + 1));
    DFB_TRY(total_dev.alloc(ctx, 1));
    {
        KTimer t(ctx, DFB_K_REGIONS);
        if (gflag)
            count_starts_flagged_kernel<<<flag_grid, 256, 0, ctx->stream>>>(mask, seg, rows, W, Gr, counts.p, gflag);
        else
            count_starts_group_kernel<<<group_grid, 256, 0, ctx->stream>>>(mask, seg, rows, W, Gr, counts.p, gflag);
    }
    DFB_TRY(exclusive_scan_u32(ctx, counts.p, offsets.p, groups, total_dev.p, DFB_K_REGIONS));
    // per-row counts: offsets at the row starts (strided gather) + total, read back through pinned
    // memory with one synchronisation (which also publishes a deferred K2 cursor, if any)
    int64_t* row_off = (int64_t*)arena_alloc(ctx, ((size_t)rows + 1) * sizeof(int64_t));
    DFB_REQUIRE(row_off != nullptr, "pinned arena allocation failed");
    DFB_CUDA(cudaMemcpy2DAsync(row_off, sizeof(int64_t), offsets.p, (size_t)Gr * sizeof(int64_t), sizeof(int64_t),
                               (size_t)rows, cudaMemcpyDeviceToHost, ctx->stream));
    DFB_CUDA(cudaMemcpyAsync(row_off + rows, total_dev.p, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    DFB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (compact && compact->used_host && *compact->used_host > compact->capacity) {
        out->total = -1;  // exceedance values overflowed: the caller re-runs K2 with the exact capacity
        return DFB_OK;
    }
    for (int64_t r = 0; r < rows; ++r) out->row_counts[(size_t)r] = row_off[(size_t)r + 1] - row_off[(size_t)r];
    out->total = row_off[(size_t)rows];
    if (out->total == 0) return DFB_OK;
    DFB_REQUIRE(out->total <= INT32_MAX, "more than 2^31-1 regions");
    DFB_TRY(out->area.alloc(ctx, (size_t)out->total));
    DFB_TRY(out->stat.alloc(ctx, (size_t)out->total));
    DFB_TRY(out->width.alloc(ctx, (size_t)out->total));
    if (want_index) {
        DFB_TRY(out->istart.alloc(ctx, (size_t)out->total));
        DFB_TRY(out->iend.alloc(ctx, (size_t)out->total));
    }
    if (want_perm) DFB_TRY(out->perm.alloc(ctx, (size_t)out->total));
    EmitParams P;
    memset(&P, 0, sizeof P);
    P.mask = mask;
    P.seg = seg;
    P.rows = rows;
    P.W = W;
    P.N = N;
    P.sides = sides;
    P.dense = dense;
    P.dense_stride = dense_stride;
    if (compact) {
        P.off = compact->off.p;
        P.tiles = compact->tiles;
        P.tile_bases = compact->tile_bases;
        P.vals = compact->vals.p;
    }
    P.area = out->area.p;
    P.stat = out->stat.p;
    P.width = out->width.p;
    P.istart = out->istart.p;
    P.iend = out->iend.p;
    P.perm = out->perm.p;
    DevBuf<int64_t> start_pos;
    DFB_TRY(start_pos.alloc(ctx, (size_t)out->total));
    DevBuf<int32_t> long_list;  // regions handed from the thread walk to the warp walk
    DevBuf<unsigned int> long_count;
    DFB_TRY(long_list.alloc(ctx, (size_t)out->total));
    DFB_TRY(long_count.alloc(ctx, 1));
    DFB_TRY(long_count.zero());
    {
        KTimer t(ctx, DFB_K_REGIONS);
        if (gflag)
            fill_starts_flagged_kernel<<<flag_grid, 256, 0, ctx->stream>>>(mask, seg, offsets.p, rows, W, Gr, start_pos.p,
                                                                           gflag);
        else
            fill_starts_group_kernel<<<group_grid, 256, 0, ctx->stream>>>(mask, seg, offsets.p, rows, W, Gr, start_pos.p,
                                                                          gflag);
    }
    {
        KTimer t(ctx, DFB_K_REGIONS);
        const unsigned grid = (unsigned)ceil_div(out->total, 128);
        if (compact)
            walk_regions_kernel<true><<<grid, 128, 0, ctx->stream>>>(P, start_pos.p, out->total, long_list.p, long_count.p);
        else
            walk_regions_kernel<false><<<grid, 128, 0, ctx->stream>>>(P, start_pos.p, out->total, long_list.p,
                                                                      long_count.p);
    }
    {
        KTimer t(ctx, DFB_K_REGIONS);
        const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(out->total, 4), (int64_t)ctx->sm_count * 16);
        if (compact) walk_long_regions_kernel<true><<<grid, 128, 0, ctx->stream>>>(P, start_pos.p, long_list.p, long_count.p);
        else walk_long_regions_kernel<false><<<grid, 128, 0, ctx->stream>>>(P, start_pos.p, long_list.p, long_count.p);
    }
    DFB_CUDA(cudaGetLastError());
    return DFB_OK;
}

// ------------------------------------------------------------------------------------------
// genomic coordinates and clusters (full mode)

__device__ __forceinline__ int32_t genomic_pos(const int32_t* gstart, const int64_t* cum, int64_t K, int64_t idx0) {
    // largest k with cum[k] <= idx0
    int64_t lo = 0, hi = K;  // cum has K+1 entries, cum[K] = N > idx0
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (cum[mid] <= idx0) lo = mid;
        else hi = mid;
    }
    return (int32_t)((int64_t)gstart[lo] + (idx0 - cum[lo]));
}

__global__ void genomic_map_kernel(const int32_t* __restrict__ istart, const int32_t* __restrict__ iend, int64_t R,
                                   const int32_t* __restrict__ gstart, const int64_t* __restrict__ cum, int64_t K,
                                   int32_t* __restrict__ gs, int32_t* __restrict__ ge) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    gs[r] = genomic_pos(gstart, cum, K, (int64_t)istart[r] - 1);
    ge[r] = genomic_pos(gstart, cum, K, (int64_t)iend[r] - 1);
}

// flags / widths feeding the scans
__global__ void cluster_prep_kernel(const int32_t* __restrict__ gs, const int32_t* __restrict__ ge,
                                    const int32_t* __restrict__ istart, const int32_t* __restrict__ iend, int64_t R,
                                    int32_t max_cluster_gap, uint32_t* __restrict__ flag, uint32_t* __restrict__ gw,
                                    uint32_t* __restrict__ iw) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    bool open = true;
    if (r > 0) open = ((int64_t)gs[r] - (int64_t)ge[r - 1] - 1) > (int64_t)max_cluster_gap;
    flag[r] = open ? 1u : 0u;
    gw[r] = (uint32_t)(ge[r] - gs[r] + 1);
    iw[r] = (uint32_t)(iend[r] - istart[r] + 1);
}

__global__ void cluster_prod_kernel(const int64_t* __restrict__ flag_excl, const uint32_t* __restrict__ flag,
                                    const uint32_t* __restrict__ gw, int64_t R, int64_t* __restrict__ rc,
                                    int64_t* __restrict__ prod) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const int64_t c = flag_excl[r] + flag[r];  // id of the merged covered range holding region r
    rc[r] = c;
    prod[r] = c * (int64_t)gw[r];
}

// sum of the first t entries of the per-covered-base cluster-id vector
__device__ __forceinline__ int64_t cluster_prefix(const int64_t* G, const int64_t* Phi, const int64_t* rc, int64_t R,
                                                  int64_t gtot, int64_t t) {
    if (t > gtot) t = gtot;
    int64_t lo = 0, hi = R;  // largest j < R with G[j] <= t
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (G[mid] <= t) lo = mid;
        else hi = mid;
    }
    return Phi[lo] + (t - G[lo]) * rc[lo];
}

// clusterFinal = as.integer(mean(Views(cluster, IRanges(cumulative INDEX widths)))) (:259-265)
__global__ void cluster_final_kernel(const int64_t* __restrict__ G, const int64_t* __restrict__ Phi,
                                     const int64_t* __restrict__ rc, const int64_t* __restrict__ CW,
                                     const uint32_t* __restrict__ iw, const int64_t* __restrict__ gtot, int64_t R,
                                     int32_t* __restrict__ cf, uint32_t* __restrict__ gflag) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const int64_t t0 = CW[r], t1 = CW[r] + (int64_t)iw[r];
    const int64_t s = cluster_prefix(G, Phi, rc, R, *gtot, t1) - cluster_prefix(G, Phi, rc, R, *gtot, t0);
    cf[r] = (int32_t)trunc((double)s / (double)iw[r]);
}

__global__ void cluster_gflag_kernel(const int32_t* __restrict__ cf, int64_t R, uint32_t* __restrict__ gflag) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    gflag[r] = (r == 0 || cf[r] != cf[r - 1]) ? 1u : 0u;
}

__global__ void cluster_bounds_kernel(const int32_t* __restrict__ cf, const int64_t* __restrict__ gexcl,
                                      const uint32_t* __restrict__ gflag, int64_t R, int64_t* __restrict__ gfirst,
                                      int64_t* __restrict__ glast) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const int64_t grp = gexcl[r] + gflag[r] - 1;  // 0-based group (groups = runs of equal cf)
    if (gflag[r]) gfirst[grp] = r;
    if (r == R - 1 || cf[r + 1] != cf[r]) glast[grp] = r;
}

// clusterWidth[clusterFinal]: tapply result indexed by POSITION (:266-271); NaN encodes NA
__global__ void cluster_len_kernel(const int32_t* __restrict__ cf, const int32_t* __restrict__ gs,
                                   const int32_t* __restrict__ ge, const int64_t* __restrict__ gfirst,
                                   const int64_t* __restrict__ glast, const int64_t* __restrict__ ngroups, int64_t R,
                                   double* __restrict__ cl) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const int64_t c = cf[r];
    if (c >= 1 && c <= *ngroups) cl[r] = (double)((int64_t)ge[glast[c - 1]] - (int64_t)gs[gfirst[c - 1]] + 1);
    else cl[r] = nan("");
}

// The whole cluster chain of one side -- the six element-wise kernels and five scans below -- by ONE block when the
// region table is small (it nearly always is: tens of thousands of regions at most): same formulas, same scratch
// arrays, one launch instead of eleven.  Block-wide exclusive scan of get(r), r < R, into out[]; returns the total.
template <typename Get>
__device__ __forceinline__ int64_t block_scan_excl(int64_t R, Get get, int64_t* __restrict__ out) {
    constexpr int IT = 4;
    __shared__ int64_t wsum_s[32];
    __shared__ int64_t carry_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __syncthreads();  // previous users of wsum_s / carry_s and of the arrays get() reads are done
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int64_t base = 0; base < R; base += (int64_t)blockDim.x * IT) {
        int64_t v[IT];
        int64_t tsum = 0;
#pragma unroll
        for (int i = 0; i < IT; ++i) {
            const int64_t r = base + (int64_t)tid * IT + i;
            v[i] = r < R ? get(r) : 0;
            tsum += v[i];
        }
        int64_t inc = tsum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int64_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) wsum_s[warp] = inc;
        __syncthreads();
        int64_t wprefix = 0;
        for (int w = 0; w < warp; ++w) wprefix += wsum_s[w];
        int64_t excl = carry_s + wprefix + inc - tsum;
#pragma unroll
        for (int i = 0; i < IT; ++i) {
            const int64_t r = base + (int64_t)tid * IT + i;
            if (r < R) out[r] = excl;
            excl += v[i];
        }
        __syncthreads();
        if (tid == (int)blockDim.x - 1) carry_s = excl;
        __syncthreads();
    }
    return carry_s;
}

__global__ void __launch_bounds__(1024) cluster_one_block_kernel(const int32_t* __restrict__ gs, const int32_t* __restrict__ ge,
                                                                 const int32_t* __restrict__ is, const int32_t* __restrict__ ie,
                                                                 int64_t R, int32_t max_cluster_gap, int64_t* fexcl, int64_t* G,
                                                                 int64_t* rc, int64_t* Phi, int64_t* CW, int64_t* gexcl,
                                                                 int64_t* gfirst, int64_t* glast, int32_t* cf, double* cl) {
    auto flag = [&](int64_t r) -> int64_t {
        return (r == 0 || ((int64_t)gs[r] - (int64_t)ge[r - 1] - 1) > (int64_t)max_cluster_gap) ? 1 : 0;
    };
    auto gw = [&](int64_t r) -> int64_t { return (int64_t)(ge[r] - gs[r] + 1); };
    auto iw = [&](int64_t r) -> int64_t { return (int64_t)(ie[r] - is[r] + 1); };
    block_scan_excl(R, flag, fexcl);
    const int64_t gtot = block_scan_excl(R, gw, G);  // covered bases
    for (int64_t r = threadIdx.x; r < R; r += blockDim.x) rc[r] = fexcl[r] + flag(r);  // id of the merged covered range
    block_scan_excl(R, [&](int64_t r) -> int64_t { return rc[r] * gw(r); }, Phi);
    block_scan_excl(R, iw, CW);
    __syncthreads();
    for (int64_t r = threadIdx.x; r < R; r += blockDim.x) {  // clusterFinal (R/findRegions.R:259-265)
        const int64_t t0 = CW[r], t1 = CW[r] + iw(r);
        const int64_t sum = cluster_prefix(G, Phi, rc, R, gtot, t1) - cluster_prefix(G, Phi, rc, R, gtot, t0);
        cf[r] = (int32_t)trunc((double)sum / (double)iw(r));
    }
    const int64_t ngroups =
        block_scan_excl(R, [&](int64_t r) -> int64_t { return (r == 0 || cf[r] != cf[r - 1]) ? 1 : 0; }, gexcl);
    for (int64_t r = threadIdx.x; r < R; r += blockDim.x) {  // first / last region of every run of equal clusterFinal
        const bool first = r == 0 || cf[r] != cf[r - 1];
        const int64_t grp = gexcl[r] + (first ? 1 : 0) - 1;
        if (first) gfirst[grp] = r;
        if (r == R - 1 || cf[r + 1] != cf[r]) glast[grp] = r;
    }
    __syncthreads();
    for (int64_t r = threadIdx.x; r < R; r += blockDim.x) {  // clusterWidth[clusterFinal] (:266-271); NaN encodes NA
        const int64_t c = cf[r];
        if (c >= 1 && c <= ngroups) cl[r] = (double)((int64_t)ge[glast[c - 1]] - (int64_t)gs[gfirst[c - 1]] + 1);
        else cl[r] = nan("");
    }
}

static int cluster_side(dfb_ctx* ctx, const int32_t* gs, const int32_t* ge, const int32_t* is, const int32_t* ie,
                        int64_t R, int32_t max_cluster_gap, int32_t* cf, double* cl) {
    if (R == 0) return DFB_OK;
    // (up to 4096 regions: beyond that the block's serial share of the binary searches of clusterFinal costs more
    // than the ten launches it saves -- 0.22 ms against 0.1 ms at 22 000 regions, visible end to end at C2)
    if (R <= 4096 && ctx->opt_cluster_one_block) {
        DevBuf<int64_t> scratch;  // fexcl, G, rc, Phi, CW, gexcl, gfirst, glast
        DFB_TRY(scratch.alloc(ctx, (size_t)8 * R));
        int64_t* a = scratch.p;
        {
            KTimer t(ctx, DFB_K_REGIONS);
            cluster_one_block_kernel<<<1, 1024, 0, ctx->stream>>>(gs, ge, is, ie, R, max_cluster_gap, a, a + R, a + 2 * R, a + 3 * R,
                                                                  a + 4 * R, a + 5 * R, a + 6 * R, a + 7 * R, cf, cl);
        }
        DFB_CUDA(cudaGetLastError());
        return DFB_OK;
    }
    const unsigned grid = (unsigned)ceil_div(R, 256);
    DevBuf<uint32_t> flag, gw, iw, gflag;
    DevBuf<int64_t> fexcl, G, rc, prod, Phi, CW, gexcl, gfirst, glast, tot;
    DFB_TRY(flag.alloc(ctx, R));
    DFB_TRY(gw.alloc(ctx, R));
    DFB_TRY(iw.alloc(ctx, R));
    DFB_TRY(gflag.alloc(ctx, R));
    DFB_TRY(fexcl.alloc(ctx, R));
    DFB_TRY(G.alloc(ctx, R));
    DFB_TRY(rc.alloc(ctx, R));
    DFB_TRY(prod.alloc(ctx, R));
    DFB_TRY(Phi.alloc(ctx, R));
    DFB_TRY(CW.alloc(ctx, R));
    DFB_TRY(gexcl.alloc(ctx, R));
    DFB_TRY(gfirst.alloc(ctx, R));
    DFB_TRY(glast.alloc(ctx, R));
    DFB_TRY(tot.alloc(ctx, 4));
    ctx->launches += 6;  // the six element-wise kernels below (scans count themselves)
    ctx->klaunch[DFB_K_REGIONS] += 6;
    cluster_prep_kernel<<<grid, 256, 0, ctx->stream>>>(gs, ge, is, ie, R, max_cluster_gap, flag.p, gw.p, iw.p);
    DFB_TRY(exclusive_scan_u32(ctx, flag.p, fexcl.p, R, tot.p + 0, DFB_K_REGIONS));
    DFB_TRY(exclusive_scan_u32(ctx, gw.p, G.p, R, tot.p + 1, DFB_K_REGIONS));  // tot[1] = covered bases
    cluster_prod_kernel<<<grid, 256, 0, ctx->stream>>>(fexcl.p, flag.p, gw.p, R, rc.p, prod.p);
    DFB_TRY(exclusive_scan_i64(ctx, prod.p, Phi.p, R, tot.p + 2, DFB_K_REGIONS));
    DFB_TRY(exclusive_scan_u32(ctx, iw.p, CW.p, R, tot.p + 2, DFB_K_REGIONS));
    cluster_final_kernel<<<grid, 256, 0, ctx->stream>>>(G.p, Phi.p, rc.p, CW.p, iw.p, tot.p + 1, R, cf, gflag.p);
    cluster_gflag_kernel<<<grid, 256, 0, ctx->stream>>>(cf, R, gflag.p);
    DFB_TRY(exclusive_scan_u32(ctx, gflag.p, gexcl.p, R, tot.p + 3, DFB_K_REGIONS));  // tot[3] = #groups
    cluster_bounds_kernel<<<grid, 256, 0, ctx->stream>>>(cf, gexcl.p, gflag.p, R, gfirst.p, glast.p);
    cluster_len_kernel<<<grid, 256, 0, ctx->stream>>>(cf, gs, ge, gfirst.p, glast.p, tot.p + 3, R, cl);
    DFB_CUDA(cudaGetLastError());
    return DFB_OK;
}

// ------------------------------------------------------------------------------------------
// mean(Views(v, IRanges(indexStart, indexEnd))) -- run-by-run like the areas

__global__ void view_means_kernel(const double* __restrict__ v, const int32_t* __restrict__ is,
                                  const int32_t* __restrict__ ie, int64_t R, double* __restrict__ out) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    int64_t i = (int64_t)is[r] - 1;
    const int64_t e = ie[r];
    double acc = 0.0;
    double cur = v[i];
    double runlen = 1.0;
    for (int64_t j = i + 1; j < e; ++j) {
        const double x = v[j];
        if (x == cur) {
            runlen += 1.0;
        } else {
            acc = __dadd_rn(acc, __dmul_rn(cur, runlen));
            cur = x;
            runlen = 1.0;
        }
    }
    acc = __dadd_rn(acc, __dmul_rn(cur, runlen));
    out[r] = __ddiv_rn(acc, (double)(e - i));
}

// ------------------------------------------------------------------------------------------
// observed regions from a dense device vector

int find_regions_dense(dfb_ctx* ctx, const double* x_dev, const PositionIndex& pos, double lo, double hi,
                       int32_t max_region_gap, int32_t max_cluster_gap, bool basic, dfb_regions** out, bool defer) {
    *out = nullptr;
    const int64_t N = pos.N;
    if (N == 0 || pos.K == 0) return DFB_ENOREGIONS;  // warning("Found no regions") :162-165
    const int64_t W = ceil_div(N, 32);
    DevBuf<uint32_t> seg, mask;
    DFB_TRY(build_segment_starts(ctx, pos, max_region_gap, &seg));
    DFB_TRY(mask.alloc(ctx, (size_t)2 * W));
    {
        KTimer t(ctx, DFB_K_REGIONS);
        threshold_mask_kernel<<<grid_for(W * 32, 256, ctx->sm_count), 256, 0, ctx->stream>>>(x_dev, N, hi, lo, mask.p,
                                                                                           mask.p + W, W);
    }
    RegionSet rs;
    // both rows read the same dense vector (stride 0)
    DFB_TRY(find_regions_rows(ctx, N, 2, 1, mask.p, seg.p, x_dev, 0, nullptr, true, false, &rs));
    if (rs.total == 0) return DFB_ENOREGIONS;
    dfb_regions* r = new dfb_regions();
    r->ctx = ctx;
    r->R = rs.total;
    r->R_up = rs.row_counts[0];
    r->basic = basic;
    const size_t R = (size_t)rs.total;
    DevBuf<int32_t> gs, ge, cf;
    DevBuf<double> cl;
    if (!basic) {
        int rc0 = gs.alloc(ctx, R);
        if (rc0 >= 0) rc0 = ge.alloc(ctx, R);
        if (rc0 >= 0) rc0 = cf.alloc(ctx, R);
        if (rc0 >= 0) rc0 = cl.alloc(ctx, R);
        if (rc0 < 0) {
            delete r;
            return rc0;
        }
        {
            KTimer t(ctx, DFB_K_REGIONS);
            genomic_map_kernel<<<(unsigned)ceil_div((int64_t)R, 256), 256, 0, ctx->stream>>>(
                rs.istart.p, rs.iend.p, (int64_t)R, pos.run_gstart.p, pos.run_cum.p, pos.K, gs.p, ge.p);
        }
        // clusters are computed separately for the up and the dn set (loop at :231)
        int rc1 = cluster_side(ctx, gs.p, ge.p, rs.istart.p, rs.iend.p, r->R_up, max_cluster_gap, cf.p, cl.p);
        int rc2 = rc1 < 0 ? rc1
                          : cluster_side(ctx, gs.p + r->R_up, ge.p + r->R_up, rs.istart.p + r->R_up,
                                         rs.iend.p + r->R_up, r->R - r->R_up, max_cluster_gap, cf.p + r->R_up,
                                         cl.p + r->R_up);
        if (rc2 < 0) {
            delete r;
            return rc2;
        }
    }
    // results travel through pinned memory: all copies are asynchronous, one synchronisation
    const int32_t *h_gs = nullptr, *h_ge = nullptr, *h_cf = nullptr;
    const double* h_cl = nullptr;
    if (!basic) {
        h_gs = d2h_async_pinned(ctx, gs.p, R);
        h_ge = d2h_async_pinned(ctx, ge.p, R);
        h_cf = d2h_async_pinned(ctx, cf.p, R);
        h_cl = d2h_async_pinned(ctx, cl.p, R);
    }
    const double* h_val = d2h_async_pinned(ctx, rs.stat.p, R);
    const double* h_area = d2h_async_pinned(ctx, rs.area.p, R);
    const int32_t* h_is = d2h_async_pinned(ctx, rs.istart.p, R);
    const int32_t* h_ie = d2h_async_pinned(ctx, rs.iend.p, R);
    if (!h_val || !h_area || !h_is || !h_ie || (!basic && (!h_gs || !h_ge || !h_cf || !h_cl))) {
        set_error("find_regions: pinned read-back failed");
        delete r;
        return DFB_ECUDA;
    }
    r->pend = {h_gs, h_ge, h_cf, h_cl, h_val, h_area, h_is, h_ie, true, ctx->stream};
    r->d_area = std::move(rs.area);
    r->d_is = std::move(rs.istart);
    r->d_ie = std::move(rs.iend);
    if (!defer) {
        int rc = finalize_regions(ctx, r);
        if (rc < 0) {
            delete r;
            return rc;
        }
    }
    *out = r;
    return DFB_OK;
}

// Second half of find_regions_dense: wait for the asynchronous read-backs and fill the host
// vectors.  dfb_calculate_pvalues defers it until the permutation scan has been enqueued, so the
// device never idles while the host copies region tables around.
int finalize_regions(dfb_ctx* ctx, dfb_regions* r) {
    if (!r->pend.active) return DFB_OK;
    (void)ctx;
    cudaError_t e = cudaStreamSynchronize(r->pend.stream);
    if (e != cudaSuccess) {
        set_error("find_regions: %s", cudaGetErrorString(e));
        return DFB_ECUDA;
    }
    const size_t R = (size_t)r->R;
    const auto& p = r->pend;
    r->value.assign(p.val, p.val + R);
    r->area.assign(p.area, p.area + R);
    r->index_start.assign(p.is, p.is + R);
    r->index_end.assign(p.ie, p.ie + R);
    if (!r->basic) {
        r->start.assign(p.gs, p.gs + R);
        r->end.assign(p.ge, p.ge + R);
        r->cluster.assign(p.cf, p.cf + R);
        r->cluster_l.assign(p.cl, p.cl + R);
    }
    r->pend.active = false;
    return DFB_OK;
}

}  // namespace dfb

using namespace dfb;

extern "C" {

int dfb_find_regions(dfb_ctx* ctx, const dfb_cov* cov, const double* x, double cutoff_lo, double cutoff_hi,
                     int32_t max_region_gap, int32_t max_cluster_gap, int basic, dfb_regions** out) {
    DFB_REQUIRE(ctx && cov && out, "dfb_find_regions: NULL argument");
    DFB_REQUIRE(cutoff_lo <= cutoff_hi, "cutoff pair must be sorted (L <= U)");
    arena_reset(ctx);
    *out = nullptr;
    const double* xd = nullptr;
    DevBuf<double> tmp;
    if (x) {
        DFB_TRY(tmp.alloc(ctx, (size_t)std::max<int64_t>(cov->N, 1)));
        DFB_TRY(h2d(ctx, tmp.p, x, (size_t)cov->N * sizeof(double)));
        xd = tmp.p;
    } else {
        if (!cov->has_f) {
            set_error("dfb_find_regions: no statistics given and none cached (run dfb_fstats first)");
            return DFB_ESTATE;
        }
        xd = cov->fstats.p;
    }
    int rc = find_regions_dense(ctx, xd, cov->pos, cutoff_lo, cutoff_hi, max_region_gap, max_cluster_gap, basic != 0,
                                out);
    cudaStreamSynchronize(ctx->stream);  // tmp must outlive the kernels reading it
    return rc;
}

int dfb_find_regions_pos(dfb_ctx* ctx, const double* x, int64_t N, const int32_t* pos_len, int64_t pos_nruns,
                         int pos_first, double cutoff_lo, double cutoff_hi, int32_t max_region_gap,
                         int32_t max_cluster_gap, int basic, dfb_regions** out) {
    DFB_REQUIRE(ctx && x && pos_len && out, "dfb_find_regions_pos: NULL argument");
    DFB_REQUIRE(cutoff_lo <= cutoff_hi, "cutoff pair must be sorted (L <= U)");
    arena_reset(ctx);
    *out = nullptr;
    PositionIndex pos;
    DFB_TRY(pos.build(ctx, pos_len, pos_nruns, pos_first));
    DFB_REQUIRE(pos.N == N, "length(fstats) = %lld does not match sum(position) = %lld", (long long)N,
                (long long)pos.N);
    DevBuf<double> tmp;
    DFB_TRY(tmp.alloc(ctx, (size_t)std::max<int64_t>(N, 1)));
    DFB_TRY(h2d(ctx, tmp.p, x, (size_t)N * sizeof(double)));
    int rc = find_regions_dense(ctx, tmp.p, pos, cutoff_lo, cutoff_hi, max_region_gap, max_cluster_gap, basic != 0,
                                out);
    cudaStreamSynchronize(ctx->stream);
    return rc;
}

void dfb_regions_free(dfb_regions* r) {
    if (!r) return;
    // some of the buffers were allocated on the auxiliary stream (the observed-region chain of
    // dfb_calculate_pvalues) and are read by launches on the main one: both must have drained before the
    // stream-ordered frees are queued
    if (r->ctx) {
        cudaStreamSynchronize(r->ctx->stream);
        if (r->ctx->aux) cudaStreamSynchronize(r->ctx->aux);
    }
    delete r;
}
int64_t dfb_regions_count(const dfb_regions* r) { return r ? r->R : 0; }
int64_t dfb_regions_count_up(const dfb_regions* r) { return r ? r->R_up : 0; }

int dfb_regions_get(const dfb_regions* r, int32_t* start, int32_t* end, double* value, double* area,
                    int32_t* index_start, int32_t* index_end, int32_t* cluster, double* cluster_l) {
    DFB_REQUIRE(r, "dfb_regions_get: NULL handle");
    const size_t R = (size_t)r->R;
    const bool ordered = !r->order.empty();
    auto put = [&](auto* dst, const auto& src) {
        if (!dst || src.size() != R) return;
        for (size_t k = 0; k < R; ++k) dst[k] = src[ordered ? (size_t)r->order[k] : k];
    };
    if ((start || end || cluster || cluster_l) && r->basic) {
        set_error("dfb_regions_get: genomic coordinates / clusters are not available in basic mode");
        return DFB_ESTATE;
    }
    put(start, r->start);
    put(end, r->end);
    put(value, r->value);
    put(area, r->area);
    put(index_start, r->index_start);
    put(index_end, r->index_end);
    put(cluster, r->cluster);
    put(cluster_l, r->cluster_l);
    return DFB_OK;
}

int dfb_regions_get_order(const dfb_regions* r, int32_t* order) {
    DFB_REQUIRE(r && order, "dfb_regions_get_order: NULL argument");
    for (int64_t k = 0; k < r->R; ++k) order[k] = r->order.empty() ? (int32_t)k : r->order[(size_t)k];
    return DFB_OK;
}

int dfb_regions_get_pvalues(const dfb_regions* r, double* pvalues) {
    DFB_REQUIRE(r && pvalues, "dfb_regions_get_pvalues: NULL argument");
    if (r->pvalues.size() != (size_t)r->R) {
        set_error("dfb_regions_get_pvalues: p-values have not been computed for this result");
        return DFB_ESTATE;
    }
    memcpy(pvalues, r->pvalues.data(), (size_t)r->R * sizeof(double));
    return DFB_OK;
}

int dfb_regions_view_means(dfb_ctx* ctx, const dfb_regions* r, const dfb_cov* cov, int which, double* out) {
    DFB_REQUIRE(ctx && r && cov && out, "dfb_regions_view_means: NULL argument");
    const double* v = nullptr;
    if (which < 0) {
        if (!cov->has_mean) {
            set_error("meanCoverage has not been computed (run dfb_preprocess or filter with returnMean)");
            return DFB_ESTATE;
        }
        v = cov->mean.p;
    } else {
        DFB_REQUIRE(which < cov->G, "group %d out of range (%d groups)", which, cov->G);
        v = cov->gmean.p + (size_t)which * cov->ld;
    }
    const int64_t R = r->R;
    if (R == 0) return DFB_OK;
    DevBuf<double> o;
    DFB_TRY(o.alloc(ctx, (size_t)R));
    {
        KTimer t(ctx, DFB_K_REGIONS);
        view_means_kernel<<<(unsigned)ceil_div(R, 128), 128, 0, ctx->stream>>>(v, r->d_is.p, r->d_ie.p, R, o.p);
    }
    DFB_CUDA(cudaGetLastError());
    std::vector<double> tmp((size_t)R);
    DFB_TRY(d2h(ctx, tmp.data(), o.p, (size_t)R * sizeof(double)));
    const bool ordered = !r->order.empty();
    for (int64_t k = 0; k < R; ++k) out[k] = tmp[ordered ? (size_t)r->order[(size_t)k] : (size_t)k];
    return DFB_OK;
}

// The same per-region means for a caller-supplied host vector (coveragePrep$meanCoverage or one of
// coveragePrep$groupMeans when the caller did not take them from this handle, R/calculatePvalues.R:222-262).
int dfb_regions_view_means_host(dfb_ctx* ctx, const dfb_regions* r, const double* v, int64_t N, double* out) {
    DFB_REQUIRE(ctx && r && v && out, "dfb_regions_view_means_host: NULL argument");
    const int64_t R = r->R;
    if (R == 0) return DFB_OK;
    for (int64_t k = 0; k < R; ++k)
        DFB_REQUIRE(r->index_end.empty() || (int64_t)r->index_end[(size_t)k] <= N,
                    "region %lld ends at index %d beyond the %lld values given", (long long)k, r->index_end[(size_t)k],
                    (long long)N);
    DevBuf<double> vd, o;
    DFB_TRY(vd.alloc(ctx, (size_t)std::max<int64_t>(N, 1)));
    DFB_TRY(h2d_big(ctx, vd.p, v, (size_t)N * sizeof(double)));
    DFB_TRY(o.alloc(ctx, (size_t)R));
    {
        KTimer t(ctx, DFB_K_REGIONS);
        view_means_kernel<<<(unsigned)ceil_div(R, 128), 128, 0, ctx->stream>>>(vd.p, r->d_is.p, r->d_ie.p, R, o.p);
    }
    DFB_CUDA(cudaGetLastError());
    std::vector<double> tmp((size_t)R);
    DFB_TRY(d2h(ctx, tmp.data(), o.p, (size_t)R * sizeof(double)));
    const bool ordered = !r->order.empty();
    for (int64_t k = 0; k < R; ++k) out[k] = tmp[ordered ? (size_t)r->order[(size_t)k] : (size_t)k];
    return DFB_OK;
}

// .clusterMakerRle on run lengths (position metadata, O(#runs)); the device path uses
// segment_start_kernel / cluster_side for the same merge rule.
int64_t dfb_cluster_maker(const int32_t* pos_len, int64_t pos_nruns, int pos_first, int32_t max_gap, int64_t* counts,
                          int64_t* starts, int64_t* ends, int64_t cap) {
    if (!pos_len && pos_nruns > 0) {
        set_error("dfb_cluster_maker: NULL run lengths");
        return DFB_EINVAL;
    }
    int64_t nclus = 0, gpos = 1, prev_end = 0, csum = 0, cur = 0;
    bool have = false;
    int v = pos_first ? 1 : 0;
    auto flush = [&]() {
        if (!have) return;
        if (nclus < cap) {
            if (counts) counts[nclus] = cur;
            if (starts) starts[nclus] = csum - cur + 1;
            if (ends) ends[nclus] = csum;
        }
        ++nclus;
    };
    for (int64_t r = 0; r < pos_nruns; ++r, v ^= 1) {
        const int64_t l = pos_len[r];
        if (l <= 0) continue;
        if (v) {
            if (have && gpos - prev_end - 1 > (int64_t)max_gap) {
                flush();
                cur = 0;
            }
            have = true;
            cur += l;
            csum += l;
            prev_end = gpos + l - 1;
        }
        gpos += l;
    }
    flush();
    return nclus;
}

}  // extern "C"
