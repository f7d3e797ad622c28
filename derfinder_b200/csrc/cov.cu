// K1 -- coverage preprocessing: Rle expansion, library-size normalisation, per-base filter,
// compaction to the dense sample-major matrix, means and the log2(x + scalefac) transform.
//
// Reference: filterData (R/filterData.R:89-241) and preprocessCoverage (R/preprocessCoverage.R:
// 97-260).  In R every step materialises Rle objects per sample; here the chromosome is cut into
// tiles of 2048 positions.  Integer coverage without a divisor is filtered in RUN space (the kept mask from
// run events and a prefix sum, the compaction in rank space: filter_events_kernel / filter_compact_events_kernel);
// doubles and real divisors go through the tile expansion kernel (a tile whose samples are all constant -- ~90% of a
// genome -- is decided with one predicate, only tiles with kept bases are expanded).  preprocessCoverage of integer
// coverage does not write the log2 matrix at all (dfb_cov::y_lazy).
#include <math.h>

#include <algorithm>
#include <memory>
#include <type_traits>

#include "data.cuh"

namespace dfb {

constexpr int K1_THREADS = 256;
constexpr int K1_PER = 8;                        // consecutive positions per thread
constexpr int K1_TILE = K1_THREADS * K1_PER;     // 2048 positions per tile

struct RleDev {
    const void* values;      // concatenated run values of all samples (int32 or double)
    const int64_t* cum;      // [total_runs + 1] global exclusive prefix of the run lengths
    const int64_t* run_off;  // [n + 1] first run of each sample in the concatenation
    int n;
    int64_t len;  // positions per sample
};

struct FilterParams {
    RleDev rle;
    const double* mapped;  // per-sample divisor or nullptr
    int filter;            // DFB_FILTER_*
    double cutoff;
    uint32_t* mask;        // [ceil(len/32)]
    uint32_t* tile_count;  // [tiles]
    const int64_t* tile_off;  // pass 1
    const int64_t* tile_run;  // [tiles][n] run holding each tile's first position (tile_runs_kernel)
    int64_t ld;
    void* out;             // [n][ld] int32 or double
    int out_is_double;
    double* mean;          // [N]
    double* mean_tmp;      // [tiles][K1_TILE] means of a tile's kept positions in tile-local rank order (events path)
    int64_t tiles;
};

// first run g in [lo, hi] (global run indices of one sample) whose end (cum[g+1]-cb) exceeds pos
__device__ __forceinline__ int64_t find_run(const int64_t* __restrict__ cum, int64_t cb, int64_t lo, int64_t hi,
                                            int64_t pos) {
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (cum[mid + 1] - cb > pos) hi = mid;
        else lo = mid + 1;
    }
    return lo;
}

// run of every sample that contains the first position of every tile: tab[tile * n + s].  All
// tiles x samples searches run in parallel here, so the tile loop of the filter kernel starts from a
// table look-up instead of 2 x log2(#runs) dependent loads per sample.
__global__ void __launch_bounds__(256) tile_runs_kernel(RleDev R, int64_t tiles, int64_t* __restrict__ tab,
                                                        int tile_len = K1_TILE) {
    const int n = R.n;
    const int64_t total = tiles * n;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t tile = i / n;
        const int s = (int)(i - tile * n);
        const int64_t lo = R.run_off[s], hi = R.run_off[s + 1] - 1;
        tab[i] = find_run(R.cum, R.cum[lo], lo, hi, tile * tile_len);
    }
}

// NORM: the library-size divisor (mapped) is present.  The division is then done ONCE PER RUN while
// the tile is expanded (a run's value is the same for all its positions), not once per position.
template <typename V, int PASS, bool NORM>
__global__ void __launch_bounds__(K1_THREADS) filter_tile_kernel(const FilterParams P) {
    using S = typename std::conditional<NORM || sizeof(V) == 8, double, int32_t>::type;  // staged element
    extern __shared__ int64_t smem_k1[];
    int64_t* r0 = smem_k1;                // [n] run containing the tile start
    int64_t* r1 = smem_k1 + P.rle.n;      // [n] run containing the tile end
    int64_t* cb_s = smem_k1 + 2 * P.rle.n;  // [n] prefix-sum base of the sample (cum[run_off[s]])
    __shared__ int wsum[K1_THREADS / 32];
    __shared__ int uniform_keep;
    const int n = P.rle.n;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const V* __restrict__ values = reinterpret_cast<const V*>(P.rle.values);
    const int64_t* __restrict__ cum = P.rle.cum;

    for (int64_t tile = blockIdx.x; tile < P.tiles; tile += gridDim.x) {
        const int64_t t0 = tile * K1_TILE;
        const int64_t t1 = min(t0 + (int64_t)K1_TILE, P.rle.len);  // exclusive
        uint32_t my_count = PASS == 1 ? P.tile_count[tile] : 0u;
        if (PASS == 1 && my_count == 0) continue;  // block-uniform
        __syncthreads();                           // r0/r1/wsum reuse across tiles
        int nonuni = 0, anypass = 0;
        for (int s = tid; s < n; s += K1_THREADS) {
            const int64_t lo = P.rle.run_off[s], hi = P.rle.run_off[s + 1] - 1, cb = cum[lo];
            const int64_t a = P.tile_run[tile * n + s];
            int64_t b = hi;  // run holding the tile's last position
            if (tile + 1 < P.tiles) {
                b = P.tile_run[(tile + 1) * n + s];          // holds position t1 ...
                if (cum[b] - cb >= t1) --b;                  // ... and starts there: t1 - 1 is in the run before
            }
            r0[s] = a;
            r1[s] = b;
            cb_s[s] = cb;
            if (a != b) nonuni = 1;
            else if (PASS == 0 && P.filter == DFB_FILTER_ONE) {
                double v = (double)values[a];
                if (P.mapped) v = v / P.mapped[s];
                if (v > P.cutoff) anypass = 1;
            }
        }
        const int tile_nonuni = __syncthreads_or(nonuni);
        if (PASS == 0) {
            // constant tile: one decision for all 2048 positions
            if (!tile_nonuni) {
                int keep;
                if (P.filter == DFB_FILTER_NONE) keep = 1;
                else if (P.filter == DFB_FILTER_ONE) keep = __syncthreads_or(anypass);
                else {
                    if (tid == 0) {
                        double sum = 0.0;
                        for (int s = 0; s < n; ++s) {
                            double v = (double)values[r0[s]];
                            if (P.mapped) v = v / P.mapped[s];
                            sum = s == 0 ? v : __dadd_rn(sum, v);
                        }
                        uniform_keep = (__ddiv_rn(sum, (double)n) > P.cutoff) ? 1 : 0;
                    }
                    __syncthreads();
                    keep = uniform_keep;
                }
                const int64_t w0 = t0 >> 5, w1 = (t1 + 31) >> 5;
                for (int64_t w = w0 + tid; w < w1; w += K1_THREADS) {
                    uint32_t bits = keep ? 0xffffffffu : 0u;
                    const int64_t rem = P.rle.len - w * 32;
                    if (rem < 32) bits &= (1u << rem) - 1u;
                    P.mask[w] = bits;
                }
                if (tid == 0) P.tile_count[tile] = keep ? (uint32_t)(t1 - t0) : 0u;
                continue;
            }
        }
        // ---- general tile: every thread walks its 8 positions through all samples, in order
        const int64_t p0 = t0 + (int64_t)tid * K1_PER;
        uint32_t bits = 0;  // kept flags of my 8 positions
        if (PASS == 1) {
            if (p0 < t1) bits = (P.mask[p0 >> 5] >> (p0 & 31)) & 0xffu;
        }
        double sum[K1_PER];
#pragma unroll
        for (int q = 0; q < K1_PER; ++q) sum[q] = 0.0;
        int64_t dst0 = 0;
        if (PASS == 1) {
            // exclusive prefix of kept counts over the threads of the block
            const int c = __popc(bits);
            int inc = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += t;
            }
            if (lane == 31) wsum[warp] = inc;
            __syncthreads();
            int wp = 0;
            for (int w = 0; w < warp; ++w) wp += wsum[w];
            dst0 = P.tile_off[tile] + wp + inc - c;
        }
        // The tile is expanded through shared memory, SC samples at a time: warp w turns the runs of
        // sample s0 + w that overlap the tile into dense values (run bounds and values are loaded 32 runs
        // per coalesced request, every run is then filled by the whole warp); afterwards every thread
        // consumes its 8 positions of those samples, in sample order.  No per-thread binary searches or
        // dependent global loads: the searches above (r0 / r1) are the only ones.
        //
        // Zero runs are never materialised.  A warp owns one 256-position sub-tile; per sample an 8-bit
        // mask says which sub-tiles hold a non-zero run (PASS 0) or a kept position (PASS 1: every
        // sample's value is needed there, zeros included).  Only those sub-tiles are zero-filled and
        // overlaid with the non-zero runs, and only their owners consume them -- adding +0.0 to the
        // ordered sums of the others would not change them, and 0 > cutoff is false for cutoff >= 0.
        constexpr int SC = sizeof(S) == 4 ? 8 : 4;  // samples per chunk: 64 KiB of staging either way
        S* vals_s = reinterpret_cast<S*>(smem_k1 + 4 * (size_t)n);
        __shared__ unsigned need_s[8];
        __shared__ unsigned kept_sub_s;
        unsigned force_need = 0u;  // sub-tiles every sample must materialise
        if (PASS == 0) {
            if (!(P.cutoff >= 0.0) || P.filter == DFB_FILTER_NONE) force_need = 0xffu;  // zeros may pass: no skipping
        } else {
            if (tid == 0) kept_sub_s = 0u;
            __syncthreads();
            const unsigned wany = __any_sync(0xffffffffu, bits != 0u) ? 1u : 0u;
            if (lane == 0 && wany) atomicOr(&kept_sub_s, 1u << warp);
            __syncthreads();
            force_need = kept_sub_s;
        }
        for (int s0 = 0; s0 < n; s0 += SC) {
            __syncthreads();  // the previous chunk has been consumed
            for (int sl = warp; sl < SC && s0 + sl < n; sl += K1_THREADS / 32) {
                const int s = s0 + sl;
                const int64_t cb = cb_s[s];
                const int64_t ga = r0[s], gb = r1[s];
                const double d = NORM ? P.mapped[s] : 1.0;
                S* dstv = vals_s + (size_t)sl * K1_TILE;
                // pass A: which sub-tiles hold a non-zero run of this sample
                unsigned need = force_need;
                if (PASS == 0 && need != 0xffu) {
                    for (int64_t g0 = ga; g0 <= gb; g0 += 32) {
                        const int64_t g = g0 + lane;
                        unsigned cov = 0u;
                        if (g <= gb && values[g] != V(0)) {
                            const int st = (int)(max(cum[g] - cb, t0) - t0);
                            const int en = (int)(min(cum[g + 1] - cb, t1) - t0);
                            if (en > st) cov = ((2u << ((en - 1) >> 8)) - 1u) & ~((1u << (st >> 8)) - 1u);
                        }
                        need |= __reduce_or_sync(0xffffffffu, cov);
                    }
                }
                if (lane == 0) need_s[sl] = need;
                // zero-fill the needed sub-tiles, then overlay the non-zero runs
                for (unsigned m = need; m; m &= m - 1) {
                    const int sub = __ffs(m) - 1;
#pragma unroll
                    for (int i = 0; i < 256 / 32; ++i) dstv[sub * 256 + i * 32 + lane] = S(0);
                }
                __syncwarp();
                if (need) {
                    for (int64_t g0 = ga; g0 <= gb; g0 += 32) {
                        const int64_t g = g0 + lane;
                        int st = 0, en = 0;
                        S val = S(0);
                        if (g <= gb) {
                            const V raw = values[g];
                            if (raw != V(0)) {
                                st = (int)(max(cum[g] - cb, t0) - t0);
                                en = (int)(min(cum[g + 1] - cb, t1) - t0);
                                if (NORM) val = (S)__ddiv_rn((double)raw, d);
                                else val = (S)raw;
                            }
                        }
                        unsigned live = __ballot_sync(0xffffffffu, en > st);
                        while (live) {
                            const int j = __ffs(live) - 1;
                            live &= live - 1;
                            const int a = __shfl_sync(0xffffffffu, st, j), e = __shfl_sync(0xffffffffu, en, j);
                            const S v = __shfl_sync(0xffffffffu, val, j);
                            for (int pp = a + lane; pp < e; pp += 32)
                                if ((need >> (pp >> 8)) & 1u) dstv[pp] = v;
                        }
                    }
                }
            }
            __syncthreads();
            if (p0 < t1) {
                for (int sl = 0; sl < SC && s0 + sl < n; ++sl) {
                    if (!((need_s[sl] >> warp) & 1u)) continue;  // all zero here and nothing to write: no-op
                    const int s = s0 + sl;
                    const S* src = vals_s + (size_t)sl * K1_TILE + (size_t)tid * K1_PER;
                    S rawv[K1_PER];
#pragma unroll
                    for (int q = 0; q < K1_PER; ++q) rawv[q] = src[q];
                    int k = 0;
#pragma unroll
                    for (int q = 0; q < K1_PER; ++q) {
                        const int64_t pos = p0 + q;
                        if (pos < t1) {
                            const double v = (double)rawv[q];
                            if (PASS == 0) {
                                if (P.filter == DFB_FILTER_ONE) {
                                    if (v > P.cutoff) bits |= 1u << q;
                                } else {
                                    sum[q] = __dadd_rn(sum[q], v);  // sums start at +0.0: 0.0 + v == v exactly
                                }
                            } else {
                                sum[q] = __dadd_rn(sum[q], v);
                                if ((bits >> q) & 1u) {
                                    const int64_t dst = dst0 + k;
                                    if (P.out_is_double) reinterpret_cast<double*>(P.out)[(size_t)s * P.ld + dst] = v;
                                    else reinterpret_cast<int32_t*>(P.out)[(size_t)s * P.ld + dst] = (int32_t)rawv[q];
                                    ++k;
                                }
                            }
                        }
                    }
                }
            }
        }
        if (PASS == 0) {
            if (p0 < t1) {
                if (P.filter == DFB_FILTER_NONE) {
                    const int64_t cnt = min((int64_t)K1_PER, t1 - p0);
                    bits = (1u << cnt) - 1u;
                } else if (P.filter == DFB_FILTER_MEAN) {
#pragma unroll
                    for (int q = 0; q < K1_PER; ++q)
                        if (p0 + q < t1 && __ddiv_rn(sum[q], (double)n) > P.cutoff) bits |= 1u << q;
                }
            }
            // four threads share one 32-bit mask word
            uint32_t word = bits << ((tid & 3) * 8);
            word |= __shfl_xor_sync(0xffffffffu, word, 1);
            word |= __shfl_xor_sync(0xffffffffu, word, 2);
            const int64_t w = (t0 >> 5) + (tid >> 2);
            if ((tid & 3) == 0 && w * 32 < t1) P.mask[w] = word;
            int c = __popc(bits);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
            if (lane == 0) wsum[warp] = c;
            __syncthreads();
            if (tid == 0) {
                int t = 0;
                for (int w2 = 0; w2 < K1_THREADS / 32; ++w2) t += wsum[w2];
                P.tile_count[tile] = (uint32_t)t;
            }
        } else if (P.mean && p0 < t1) {
            int k = 0;
#pragma unroll
            for (int q = 0; q < K1_PER; ++q)
                if ((bits >> q) & 1u) {
                    P.mean[dst0 + k] = __ddiv_rn(sum[q], (double)n);
                    ++k;
                }
        }
    }
}

// PASS 0 in RUN space, for integer coverage without a library-size divisor ("one" and "mean" filters).
//
// The kept mask only needs, per position, the sum over the samples ("mean": sum / n > cutoff) or the number of
// samples above the cutoff ("one": count > 0).  Both are sums of INTEGERS, so they do not depend on the order of
// the additions and equal R's left-to-right double sums bit for bit (every partial sum is far below 2^53).  A
// run [st, en) of value v then is two events -- +v at st, -v at en -- in a per-tile difference array, and one
// prefix sum over the 2048 positions of the tile gives every position's sum: O(runs) work instead of the
// O(positions x samples) expansion of filter_tile_kernel (a 2048-position tile of 100 samples holds ~4 500
// runs but 204 800 position-sample pairs).  Zero runs -- ~90 % of a genome -- contribute nothing and are skipped.
// Difference array of one tile -> per-position sums.  Block-wide; on return sum of position tid * 8 + q is
// before + h[q].  (acc_s: K1_TILE + 1 entries, wsum_s: one per warp.)
template <int FILTER>
__device__ __forceinline__ void tile_event_sums(const FilterParams& P, int64_t tile, int64_t t0, int64_t t1,
                                                unsigned long long* acc_s, long long* wsum_s, long long (&h)[K1_PER],
                                                long long& before) {
    const int n = P.rle.n;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int32_t* __restrict__ values = reinterpret_cast<const int32_t*>(P.rle.values);
    const int64_t* __restrict__ cum = P.rle.cum;
    __syncthreads();  // the previous user of acc_s / wsum_s is done
    for (int i = tid; i <= K1_TILE; i += K1_THREADS) acc_s[i] = 0ull;
    __syncthreads();
    for (int s = warp; s < n; s += K1_THREADS / 32) {
        const int64_t lo = P.rle.run_off[s], hi = P.rle.run_off[s + 1] - 1, cb = cum[lo];
        const int64_t a = P.tile_run[tile * n + s];
        int64_t b = hi;  // run holding the tile's last position
        if (tile + 1 < P.tiles) {
            b = P.tile_run[(tile + 1) * n + s];
            if (cum[b] - cb >= t1) --b;  // the run holding t1 starts there: t1 - 1 is in the run before
        }
        for (int64_t g = a + lane; g <= b; g += 32) {
            const int32_t v = values[g];
            const long long c = FILTER == DFB_FILTER_MEAN ? (long long)v : (((double)v > P.cutoff) ? 1ll : 0ll);
            if (c != 0) {
                const int st = (int)(max(cum[g] - cb, t0) - t0);
                const int en = (int)(min(cum[g + 1] - cb, t1) - t0);
                atomicAdd(&acc_s[st], (unsigned long long)c);
                atomicAdd(&acc_s[en], (unsigned long long)(-c));
            }
        }
    }
    __syncthreads();
    // inclusive prefix sum over the positions: 8 per thread, then across the block
    long long run = 0;
    {  // 16-byte loads: element-wise reads at a stride of 64 bytes per thread would be bank conflicts all the way
        const ulonglong2* av = reinterpret_cast<const ulonglong2*>(acc_s) + tid * (K1_PER / 2);
#pragma unroll
        for (int q = 0; q < K1_PER; q += 2) {
            const ulonglong2 v2 = av[q >> 1];
            run += (long long)v2.x;
            h[q] = run;
            run += (long long)v2.y;
            h[q + 1] = run;
        }
    }
    long long inc = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const long long t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) wsum_s[warp] = inc;
    __syncthreads();
    before = inc - run;
    for (int w = 0; w < warp; ++w) before += wsum_s[w];
}

template <int FILTER>
__global__ void __launch_bounds__(K1_THREADS) filter_events_kernel(const FilterParams P) {
    __shared__ __align__(16) unsigned long long acc_s[K1_TILE + 2];
    __shared__ long long wsum_s[K1_THREADS / 32];
    __shared__ int cnt_s[K1_THREADS / 32];
    const int n = P.rle.n;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int64_t tile = blockIdx.x; tile < P.tiles; tile += gridDim.x) {
        const int64_t t0 = tile * K1_TILE;
        const int64_t t1 = min(t0 + (int64_t)K1_TILE, P.rle.len);  // exclusive
        long long h[K1_PER], before;
        tile_event_sums<FILTER>(P, tile, t0, t1, acc_s, wsum_s, h, before);
        const int64_t p0 = t0 + (int64_t)tid * K1_PER;
        uint32_t bits = 0;
#pragma unroll
        for (int q = 0; q < K1_PER; ++q) {
            if (p0 + q < t1) {
                const long long sum = before + h[q];
                const bool keep = FILTER == DFB_FILTER_MEAN ? (__ddiv_rn((double)sum, (double)n) > P.cutoff) : (sum > 0);
                if (keep) bits |= 1u << q;
            }
        }
        // four threads share one 32-bit mask word
        uint32_t word = bits << ((tid & 3) * 8);
        word |= __shfl_xor_sync(0xffffffffu, word, 1);
        word |= __shfl_xor_sync(0xffffffffu, word, 2);
        const int64_t w = (t0 >> 5) + (tid >> 2);
        if ((tid & 3) == 0 && w * 32 < t1) P.mask[w] = word;
        const int c = __popc(bits);
        int inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) cnt_s[warp] = inc;
        __syncthreads();
        if (tid == 0) {
            int t = 0;
            for (int w2 = 0; w2 < K1_THREADS / 32; ++w2) t += cnt_s[w2];
            P.tile_count[tile] = (uint32_t)t;
        }
        if (FILTER == DFB_FILTER_MEAN && P.mean_tmp) {
            // the sums are at hand: the means of the kept positions wait in tile-local rank order for the compaction
            int r = inc - c;
            for (int w2 = 0; w2 < warp; ++w2) r += cnt_s[w2];
            double* mt = P.mean_tmp + (size_t)tile * K1_TILE;
#pragma unroll
            for (int q = 0; q < K1_PER; ++q)
                if ((bits >> q) & 1u) mt[r++] = __ddiv_rn((double)(before + h[q]), (double)n);
        }
    }
}

// PASS 1 for the same case (integer coverage, no divisor): compaction in RANK space.  Only the kept positions of a
// tile are ever materialised: rank_s[p] = number of kept positions of the tile before p; a warp takes a sample,
// every run that holds kept positions is written by the whole warp straight to the ranks of those positions in the
// output (contiguous, so the stores coalesce) -- no block barrier inside the sample loop, nothing is expanded for
// positions that are filtered away (the expansion kernel materialises whole 256-position sub-tiles).
// meanCoverage comes from the same integer event sums as the mask (exact, so equal to R's left-to-right double sum).
__global__ void __launch_bounds__(K1_THREADS) filter_compact_events_kernel(const FilterParams P) {
    __shared__ __align__(16) unsigned long long acc_s[K1_TILE + 2];
    __shared__ long long wsum_s[K1_THREADS / 32];
    __shared__ __align__(16) int rank_s[K1_TILE + 4];
    __shared__ int wcnt_s[K1_THREADS / 32];
    extern __shared__ int64_t compact_dyn[];  // ra[n], rb[n], cb[n]
    const int n = P.rle.n;
    int64_t* ra_s = compact_dyn;
    int64_t* rb_s = compact_dyn + n;
    int64_t* cb_s = compact_dyn + 2 * (size_t)n;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int32_t* __restrict__ values = reinterpret_cast<const int32_t*>(P.rle.values);
    const int64_t* __restrict__ cum = P.rle.cum;
    for (int64_t tile = blockIdx.x; tile < P.tiles; tile += gridDim.x) {
        const int cnt = (int)P.tile_count[tile];
        if (cnt == 0) continue;  // block-uniform
        const int64_t t0 = tile * K1_TILE;
        const int64_t t1 = min(t0 + (int64_t)K1_TILE, P.rle.len);
        const int64_t p0 = t0 + (int64_t)tid * K1_PER;
        const int64_t dst_tile = P.tile_off[tile];
        uint32_t bits = 0;
        if (p0 < t1) bits = (P.mask[p0 >> 5] >> (p0 & 31)) & 0xffu;
        long long h[K1_PER], before = 0;
        const bool sums_here = P.mean && !P.mean_tmp;  // "one" filter: the mask pass counted samples, not coverage
        if (sums_here) tile_event_sums<DFB_FILTER_MEAN>(P, tile, t0, t1, acc_s, wsum_s, h, before);
        else __syncthreads();  // rank_s / wcnt_s / run tables of the previous tile have been read
        if (P.mean && P.mean_tmp)
            for (int i = tid; i < cnt; i += K1_THREADS) P.mean[dst_tile + i] = P.mean_tmp[(size_t)tile * K1_TILE + i];
        // first and last run of every sample in this tile (all samples in parallel: the per-sample loop below
        // starts from shared memory instead of a chain of dependent global loads)
        for (int s = tid; s < n; s += K1_THREADS) {
            const int64_t lo = P.rle.run_off[s], hi = P.rle.run_off[s + 1] - 1, cb = cum[lo];
            int64_t b = hi;
            if (tile + 1 < P.tiles) {
                b = P.tile_run[(tile + 1) * n + s];
                if (cum[b] - cb >= t1) --b;
            }
            ra_s[s] = P.tile_run[tile * n + s];
            rb_s[s] = b;
            cb_s[s] = cb;
        }
        // ranks of my 8 positions
        const int c = __popc(bits);
        int inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) wcnt_s[warp] = inc;
        __syncthreads();
        int r = inc - c;
        for (int w = 0; w < warp; ++w) r += wcnt_s[w];
        int rk[K1_PER];
#pragma unroll
        for (int q = 0; q < K1_PER; ++q) {
            rk[q] = r;
            if ((bits >> q) & 1u) {
                if (sums_here) P.mean[dst_tile + r] = __ddiv_rn((double)(before + h[q]), (double)n);
                ++r;
            }
        }
        reinterpret_cast<int4*>(rank_s)[tid * 2] = make_int4(rk[0], rk[1], rk[2], rk[3]);  // 16-byte stores: no bank conflicts
        reinterpret_cast<int4*>(rank_s)[tid * 2 + 1] = make_int4(rk[4], rk[5], rk[6], rk[7]);
        if (tid == K1_THREADS - 1) rank_s[K1_TILE] = r;  // == cnt
        __syncthreads();
        for (int s = warp; s < n; s += K1_THREADS / 32) {
            const int64_t a = ra_s[s], b = rb_s[s], cb = cb_s[s];
            double* dstd = reinterpret_cast<double*>(P.out) + (size_t)s * P.ld + dst_tile;
            int32_t* dsti = reinterpret_cast<int32_t*>(P.out) + (size_t)s * P.ld + dst_tile;
            for (int64_t g0 = a; g0 <= b; g0 += 32) {
                const int64_t g = g0 + lane;
                int r0 = 0, r1 = 0;
                int32_t v = 0;
                if (g <= b) {
                    const int st = (int)(max(cum[g] - cb, t0) - t0);
                    const int en = (int)(min(cum[g + 1] - cb, t1) - t0);
                    r0 = rank_s[st];
                    r1 = rank_s[en];
                    if (r1 > r0) v = values[g];
                }
                // every run with kept positions is written by the whole warp: its ranks are contiguous in the output
                unsigned live = __ballot_sync(0xffffffffu, r1 > r0);
                while (live) {
                    const int j = __ffs(live) - 1;
                    live &= live - 1;
                    const int b0 = __shfl_sync(0xffffffffu, r0, j), b1 = __shfl_sync(0xffffffffu, r1, j);
                    const int32_t bv = __shfl_sync(0xffffffffu, v, j);
                    if (P.out_is_double) {
                        const double dv = (double)bv;
                        for (int i = b0 + lane; i < b1; i += 32) dstd[i] = dv;
                    } else {
                        for (int i = b0 + lane; i < b1; i += 32) dsti[i] = bv;
                    }
                }
            }
        }
    }
}

// mask words -> run boundaries: a position starts a run when its bit differs from the previous one
__global__ void __launch_bounds__(256) mask_run_count_kernel(const uint32_t* __restrict__ mask, int64_t W, int64_t len,
                                                             uint32_t* __restrict__ counts) {
    for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < W; w += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t m = mask[w];
        const uint32_t prev = w > 0 ? (mask[w - 1] >> 31) : (~m & 1u);  // position 0 always starts a run
        uint32_t d = m ^ ((m << 1) | prev);
        const int64_t rem = len - w * 32;
        if (rem < 32) d &= (1u << rem) - 1u;
        counts[w] = __popc(d);
    }
}

__global__ void __launch_bounds__(256) mask_run_emit_kernel(const uint32_t* __restrict__ mask, int64_t W, int64_t len,
                                                            const int64_t* __restrict__ offsets,
                                                            int32_t* __restrict__ run_start) {
    for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < W; w += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t m = mask[w];
        const uint32_t prev = w > 0 ? (mask[w - 1] >> 31) : (~m & 1u);
        uint32_t d = m ^ ((m << 1) | prev);
        const int64_t rem = len - w * 32;
        if (rem < 32) d &= (1u << rem) - 1u;
        int64_t o = offsets[w];
        while (d) {
            const int b = __ffs(d) - 1;
            d &= d - 1;
            run_start[o++] = (int32_t)(w * 32 + b);
        }
    }
}

// upload n Rle columns: concatenated values + global prefix of the lengths (device scan)
// bad[0] = a column holding a run length <= 0, bad[1] = a column whose lengths do not sum to len
// (-1 when fine; the smallest offending column wins)
__global__ void __launch_bounds__(256) validate_rle_kernel(const int32_t* __restrict__ lengths,
                                                           const int64_t* __restrict__ cum,
                                                           const int64_t* __restrict__ run_off, int n,
                                                           int64_t total_runs, int64_t len, int64_t* __restrict__ bad) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_runs + n;
         i += (int64_t)gridDim.x * blockDim.x) {
        if (i < total_runs) {
            if (lengths[i] <= 0) {
                int lo = 0, hi = n;  // column of run i: largest s with run_off[s] <= i
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (run_off[mid] <= i) lo = mid;
                    else hi = mid;
                }
                atomicMin((unsigned long long*)&bad[0], (unsigned long long)lo);
            }
        } else {
            const int s = (int)(i - total_runs);
            if (cum[run_off[s + 1]] - cum[run_off[s]] != len) atomicMin((unsigned long long*)&bad[1], (unsigned long long)s);
        }
    }
}

// Expansion by scatter + scan: the runs that overlap a (2048-position tile, sample) -- found through the
// tile_runs_kernel table, so no thread searches -- drop their local index at their first position of the tile in
// shared memory, an inclusive max-scan over the positions spreads it, and every thread writes its 8 positions
// as 16-byte stores.  (Its predecessor searched the run of every thread's first position: 2 x 19 dependent loads
// per 8 positions at the chr1 shape, latency-bound at 1.3 TB/s.)
template <typename V, int THREADS>
__global__ void __launch_bounds__(THREADS) expand_rle_tiles_kernel(RleDev R, int64_t tiles, const int64_t* __restrict__ tab,
                                                                   int64_t ld, V* __restrict__ out,
                                                                   const int64_t* __restrict__ bad) {
    constexpr int TILE = THREADS * K1_PER;
    // ragged or corrupt columns (flagged by validate_rle_kernel earlier in the stream) are not expanded: the tile
    // logic relies on positive run lengths that sum to R.len; the caller reports the error when it synchronises
    if (bad && (bad[0] >= 0 || bad[1] >= 0)) return;
    __shared__ __align__(16) int head_s[TILE];
    __shared__ V val_s[TILE];
    __shared__ int wmax_s[THREADS / 32];
    const int s = blockIdx.y, n = R.n;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const V* __restrict__ values = reinterpret_cast<const V*>(R.values);
    const int64_t* __restrict__ cum = R.cum;
    const int64_t lo = R.run_off[s], hi = R.run_off[s + 1] - 1, cb = cum[lo];
    (void)lo;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int64_t t0 = tile * TILE, t1 = min(t0 + (int64_t)TILE, R.len);
        const int64_t ra = tab[tile * n + s];
        int64_t rb = hi;
        if (tile + 1 < tiles) {
            rb = tab[(tile + 1) * n + s];
            if (cum[rb] - cb >= t1) --rb;  // the run holding t1 starts there: t1 - 1 is in the run before
        }
        const int nr = (int)(rb - ra + 1);  // <= TILE: every run holds at least one position of the tile
        __syncthreads();                    // previous tile consumed
#pragma unroll
        for (int q = 0; q < K1_PER; ++q) head_s[q * THREADS + tid] = 0;
        __syncthreads();
        for (int j = tid; j < nr; j += THREADS) {
            const int64_t r = ra + j;
            const int st = (int)(max(cum[r] - cb, t0) - t0);
            head_s[st] = j;
            val_s[j] = values[r];
        }
        __syncthreads();
        // inclusive max-scan of the heads: thread-local over 8 consecutive positions, then across the block
        int h[K1_PER];
        {  // two 16-byte loads per thread: element-wise reads at stride 8 would be 8-way bank conflicts
            const int4 ha = reinterpret_cast<const int4*>(head_s)[tid * 2], hb = reinterpret_cast<const int4*>(head_s)[tid * 2 + 1];
            h[0] = ha.x; h[1] = ha.y; h[2] = ha.z; h[3] = ha.w;
            h[4] = hb.x; h[5] = hb.y; h[6] = hb.z; h[7] = hb.w;
        }
        int run = 0;
#pragma unroll
        for (int q = 0; q < K1_PER; ++q) {
            run = max(run, h[q]);
            h[q] = run;
        }
        int inc = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc = max(inc, t);
        }
        if (lane == 31) wmax_s[warp] = inc;
        __syncthreads();
        int before = __shfl_up_sync(0xffffffffu, inc, 1);
        if (lane == 0) before = 0;
        for (int w = 0; w < warp; ++w) before = max(before, wmax_s[w]);
        const int64_t p0 = t0 + (int64_t)tid * K1_PER;
        if (p0 < t1) {
            V vals[K1_PER];
#pragma unroll
            for (int q = 0; q < K1_PER; ++q) vals[q] = val_s[max(h[q], before)];
            V* dst = out + (size_t)s * ld + p0;
            if (p0 + K1_PER <= t1) {
                constexpr int NV = (int)(sizeof(V) * K1_PER / 16);
#pragma unroll
                for (int q = 0; q < NV; ++q) reinterpret_cast<uint4*>(dst)[q] = reinterpret_cast<const uint4*>(vals)[q];
            } else {
                for (int e = 0; p0 + e < t1; ++e) dst[e] = vals[e];
            }
        }
    }
}

struct RleUpload {
    DevBuf<unsigned char> values;
    DevBuf<uint32_t> lengths;
    DevBuf<int64_t> cum, run_off;
    int64_t total_runs = 0;
};

// uint16 -> 32-bit columns on the device (the wire format of short-run integer coverage, see upload_rle)
__global__ void __launch_bounds__(256) widen16_kernel(const uint16_t* __restrict__ v16, const uint16_t* __restrict__ l16,
                                                      int64_t count, int32_t* __restrict__ v32, uint32_t* __restrict__ l32) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
        v32[i] = (int32_t)v16[i];
        l32[i] = (uint32_t)l16[i];
    }
}

// try16: integer columns whose values and run lengths all fit 16 bits (filtered coverage: runs of tens of bases,
// counts in the hundreds) cross PCIe as uint16 -- the narrowing rides on the copy into the pinned ring that the host
// pool makes anyway -- and are widened again on the device: half the H2D bytes of the end-to-end pass.  A value
// that does not fit abandons the attempt and the columns travel as they are.
// Upload of Rle columns in three steps, none of which synchronises with the host (unless a 16-bit attempt
// overflows): rle_alloc (device buffers + the run offsets in pinned memory), a transfer -- rle_transfer_staged
// through the pinned ring, or direct DMAs when the caller's columns are page-locked (dfb_cov_from_rle) -- and
// rle_finish (prefix sums + validation).  The validation writes the first offending column into bad_dev[0]
// (non-positive run length) / bad_dev[1] (wrong total length), both preset to -1 by the caller, who reads them back
// when it synchronises anyway.
static int rle_alloc(dfb_ctx* ctx, int n, int dtype, const void* const* values, const int32_t* const* lengths,
                     const int64_t* nruns, int64_t len, RleUpload* up, int64_t** off_out) {
    DFB_REQUIRE(n > 0 && values && lengths && nruns, "Rle input: NULL argument");
    DFB_REQUIRE(dtype == DFB_I32 || dtype == DFB_F64, "Rle input: unknown dtype %d", dtype);
    const size_t es = dtype == DFB_I32 ? 4 : 8;
    // run offsets in pinned memory: a pageable source would make the small copy wait for the stream
    int64_t* off = (int64_t*)arena_alloc(ctx, ((size_t)n + 1) * sizeof(int64_t));
    DFB_REQUIRE(off != nullptr, "pinned arena allocation failed");
    off[0] = 0;
    for (int s = 0; s < n; ++s) {
        DFB_REQUIRE(nruns[s] > 0 || len == 0, "Rle column %d is empty", s + 1);
        off[(size_t)s + 1] = off[(size_t)s] + nruns[s];
    }
    up->total_runs = off[(size_t)n];
    DFB_TRY(up->values.alloc(ctx, (size_t)std::max<int64_t>(up->total_runs, 1) * es));
    DFB_TRY(up->lengths.alloc(ctx, (size_t)std::max<int64_t>(up->total_runs, 1)));
    DFB_TRY(up->cum.alloc(ctx, (size_t)up->total_runs + 1));
    DFB_TRY(up->run_off.alloc(ctx, (size_t)n + 1));
    *off_out = off;
    return DFB_OK;
}

static int rle_transfer_staged(dfb_ctx* ctx, int n, int dtype, const void* const* values, const int32_t* const* lengths,
                               const int64_t* nruns, RleUpload* up, bool try16) {
    const size_t es = dtype == DFB_I32 ? 4 : 8;
    bool sent = false;
    // (with fewer than six copy threads -- eight ranks sharing a 16-core host -- the pass is bound by the cores, not
    // by memory or PCIe, and the pack instructions cost more than the halved stores save: measured 17.8 against
    // 15.7 ms on two threads, 5.9 against 7.3 ms on twelve)
    if (try16 && dtype == DFB_I32 && ctx->opt_wire16 && up->total_runs >= (1 << 16) && copy_threads(ctx) >= 6) {
        DevBuf<uint16_t> v16, l16;
        DFB_TRY(v16.alloc(ctx, (size_t)up->total_runs));
        DFB_TRY(l16.alloc(ctx, (size_t)up->total_runs));
        std::vector<const int32_t*> s32((size_t)n);
        bool over = false;
        for (int s = 0; s < n; ++s) s32[(size_t)s] = reinterpret_cast<const int32_t*>(values[s]);
        DFB_TRY(h2d_gather_narrow16(ctx, v16.p, s32.data(), nruns, n, false, &over));
        if (!over) {
            for (int s = 0; s < n; ++s) s32[(size_t)s] = lengths[s];
            DFB_TRY(h2d_gather_narrow16(ctx, l16.p, s32.data(), nruns, n, false, &over));
        }
        if (!over) {
            KTimer t(ctx, DFB_K_PREP);
            widen16_kernel<<<grid_for(up->total_runs, 256, ctx->sm_count, 8), 256, 0, ctx->stream>>>(
                v16.p, l16.p, up->total_runs, reinterpret_cast<int32_t*>(up->values.p), up->lengths.p);
            sent = true;
        }
    }
    if (!sent) {
        // all columns travel packed through the pinned ring: two streams of DMAs instead of 2n
        // driver-staged pageable copies
        std::vector<const void*> srcs((size_t)n);
        std::vector<size_t> nb((size_t)n);
        for (int s = 0; s < n; ++s) {
            srcs[(size_t)s] = values[s];
            nb[(size_t)s] = (size_t)nruns[s] * es;
        }
        DFB_TRY(h2d_gather(ctx, up->values.p, srcs.data(), nb.data(), n, /*sync=*/false));  // lengths follow
        for (int s = 0; s < n; ++s) {
            srcs[(size_t)s] = lengths[s];
            nb[(size_t)s] = (size_t)nruns[s] * sizeof(int32_t);
        }
        DFB_TRY(h2d_gather(ctx, up->lengths.p, srcs.data(), nb.data(), n, /*sync=*/false));
    }
    return DFB_OK;
}

// direct DMAs from page-locked caller memory on `dma` (no host pass at all)
static int rle_transfer_pinned(int n, int dtype, const void* const* values, const int32_t* const* lengths,
                               const int64_t* nruns, const int64_t* off, RleUpload* up, cudaStream_t dma) {
    const size_t es = dtype == DFB_I32 ? 4 : 8;
    for (int s = 0; s < n; ++s) {
        if (nruns[s] == 0) continue;
        DFB_CUDA(cudaMemcpyAsync(up->values.p + (size_t)off[s] * es, values[s], (size_t)nruns[s] * es, cudaMemcpyHostToDevice, dma));
        DFB_CUDA(cudaMemcpyAsync(up->lengths.p + (size_t)off[s], lengths[s], (size_t)nruns[s] * sizeof(int32_t),
                                 cudaMemcpyHostToDevice, dma));
    }
    return DFB_OK;
}

static int rle_finish(dfb_ctx* ctx, int n, int64_t len, const int64_t* off, RleUpload* up, int64_t* bad_dev) {
    DFB_CUDA(cudaMemcpyAsync(up->run_off.p, off, ((size_t)n + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
    DFB_TRY(exclusive_scan_u32(ctx, up->lengths.p, up->cum.p, up->total_runs, up->cum.p + up->total_runs, DFB_K_PREP));
    // validation on the device (catches ragged DataFrames like R would, without an O(#runs) host pass):
    // every run length positive, every column sums to `len`
    {
        KTimer t(ctx, DFB_K_PREP);
        validate_rle_kernel<<<grid_for(up->total_runs + n, 256, ctx->sm_count, 8), 256, 0, ctx->stream>>>(
            reinterpret_cast<const int32_t*>(up->lengths.p), up->cum.p, up->run_off.p, n, up->total_runs, len, bad_dev);
    }
    DFB_CUDA(cudaGetLastError());
    return DFB_OK;
}

static int upload_rle_async(dfb_ctx* ctx, int n, int dtype, const void* const* values, const int32_t* const* lengths,
                            const int64_t* nruns, int64_t len, RleUpload* up, bool try16, int64_t* bad_dev) {
    int64_t* off = nullptr;
    DFB_TRY(rle_alloc(ctx, n, dtype, values, lengths, nruns, len, up, &off));
    DFB_TRY(rle_transfer_staged(ctx, n, dtype, values, lengths, nruns, up, try16));
    return rle_finish(ctx, n, len, off, up, bad_dev);
}

static int check_rle_flags(const int64_t* hbad, int first_column, int64_t len) {
    DFB_REQUIRE(hbad[0] < 0, "Rle column %d has a non-positive run length", first_column + (int)hbad[0] + 1);
    DFB_REQUIRE(hbad[1] < 0, "Rle column %d does not have the expected length %lld", first_column + (int)hbad[1] + 1,
                (long long)len);
    return DFB_OK;
}

static int upload_rle(dfb_ctx* ctx, int n, int dtype, const void* const* values, const int32_t* const* lengths,
                      const int64_t* nruns, int64_t len, RleUpload* up, bool try16 = false) {
    DevBuf<int64_t> bad;
    DFB_TRY(bad.alloc(ctx, 2));
    DFB_CUDA(cudaMemsetAsync(bad.p, 0xff, 2 * sizeof(int64_t), ctx->stream));  // -1 = nothing wrong
    DFB_TRY(upload_rle_async(ctx, n, dtype, values, lengths, nruns, len, up, try16, bad.p));
    int64_t hbad[2] = {-1, -1};
    DFB_TRY(d2h(ctx, hbad, bad.p, sizeof hbad));  // synchronises
    return check_rle_flags(hbad, 0, len);
}

static int alloc_matrix(dfb_cov* c) {
    c->ld = round_up(std::max<int64_t>(c->N, 1), 64);
    if (c->dtype == DFB_I32) DFB_TRY(c->xi.alloc(c->ctx, (size_t)c->n * c->ld));
    else DFB_TRY(c->xd.alloc(c->ctx, (size_t)c->n * c->ld));
    return DFB_OK;
}

// filter (or plain expansion when filter == NONE) of Rle columns into a new dfb_cov
static int filter_rle_impl(dfb_ctx* ctx, int n, int dtype, const void* const* values, const int32_t* const* lengths,
                           const int64_t* nruns, int64_t len, const double* mapped, int filter, double cutoff,
                           bool want_mean, std::vector<int32_t>* new_pos_len, int* new_pos_first, dfb_cov** out) {
    RleUpload up;
    DFB_TRY(upload_rle(ctx, n, dtype, values, lengths, nruns, len, &up));
    DevBuf<double> mapped_d;
    if (mapped) {
        for (int s = 0; s < n; ++s) DFB_REQUIRE(mapped[s] != 0, "totalMapped/targetSize is zero for sample %d", s + 1);
        DFB_TRY(mapped_d.alloc(ctx, n));
        DFB_TRY(h2d(ctx, mapped_d.p, mapped, (size_t)n * sizeof(double)));
    }
    const int64_t tiles = ceil_div(len, K1_TILE);
    const int64_t W = ceil_div(len, 32);
    DevBuf<uint32_t> mask, tile_count;
    DevBuf<int64_t> tile_off, total_dev;
    DFB_TRY(mask.alloc(ctx, (size_t)std::max<int64_t>(W, 1)));
    DFB_TRY(tile_count.alloc(ctx, (size_t)std::max<int64_t>(tiles, 1)));
    DFB_TRY(tile_off.alloc(ctx, (size_t)std::max<int64_t>(tiles, 1)));
    DFB_TRY(total_dev.alloc(ctx, 1));
    FilterParams P;
    memset(&P, 0, sizeof P);
    P.rle.values = up.values.p;
    P.rle.cum = up.cum.p;
    P.rle.run_off = up.run_off.p;
    P.rle.n = n;
    P.rle.len = len;
    P.mapped = mapped ? mapped_d.p : nullptr;
    P.filter = filter;
    P.cutoff = cutoff;
    P.mask = mask.p;
    P.tile_count = tile_count.p;
    P.tiles = tiles;
    DevBuf<int64_t> tile_run;
    DFB_TRY(tile_run.alloc(ctx, (size_t)std::max<int64_t>(tiles, 1) * n));
    P.tile_run = tile_run.p;
    if (tiles > 0) {
        KTimer t(ctx, DFB_K_PREP);
        tile_runs_kernel<<<grid_for(tiles * n, 256, ctx->sm_count, 16), 256, 0, ctx->stream>>>(P.rle, tiles, tile_run.p);
    }
    const size_t smem = (size_t)4 * n * sizeof(int64_t) + (size_t)64 * 1024;  // r0 / r1 / bases + the expansion staging
    const int grid = (int)std::min<int64_t>(std::max<int64_t>(tiles, 1), (int64_t)ctx->sm_count * 16);
    auto launch_filter = [&](int pass) -> int {
        auto go = [&](auto kern) -> int {
            DFB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            KTimer t(ctx, DFB_K_PREP);
            kern<<<grid, K1_THREADS, smem, ctx->stream>>>(P);
            return DFB_OK;
        };
        // x / 1.0 == x: with totalMapped == targetSize (regionMatrix's default) every divisor is 1 and the
        // integer staging (twice the samples per chunk) serves; the output stays double
        bool norm = false;
        if (mapped)
            for (int s2 = 0; s2 < n; ++s2) norm = norm || mapped[s2] != 1.0;
        if (dtype == DFB_I32) {
            if (pass == 0 && !norm && filter != DFB_FILTER_NONE && ctx->opt_filter_events) {
                // integer sums: the mask comes from run events, no expansion (filter_events_kernel)
                KTimer t(ctx, DFB_K_PREP);
                if (filter == DFB_FILTER_MEAN) filter_events_kernel<DFB_FILTER_MEAN><<<grid, K1_THREADS, 0, ctx->stream>>>(P);
                else filter_events_kernel<DFB_FILTER_ONE><<<grid, K1_THREADS, 0, ctx->stream>>>(P);
                return DFB_OK;
            }
            if (pass == 1 && !norm && filter != DFB_FILTER_NONE && ctx->opt_filter_events) {
                const size_t hs = (size_t)3 * n * sizeof(int64_t);
                DFB_CUDA(cudaFuncSetAttribute(filter_compact_events_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hs));
                KTimer t(ctx, DFB_K_PREP);
                filter_compact_events_kernel<<<grid, K1_THREADS, hs, ctx->stream>>>(P);
                return DFB_OK;
            }
            if (pass == 0) return norm ? go(filter_tile_kernel<int32_t, 0, true>) : go(filter_tile_kernel<int32_t, 0, false>);
            return norm ? go(filter_tile_kernel<int32_t, 1, true>) : go(filter_tile_kernel<int32_t, 1, false>);
        }
        if (pass == 0) return norm ? go(filter_tile_kernel<double, 0, true>) : go(filter_tile_kernel<double, 0, false>);
        return norm ? go(filter_tile_kernel<double, 1, true>) : go(filter_tile_kernel<double, 1, false>);
    };
    DevBuf<double> mean_tmp;
    {
        bool norm0 = false;
        if (mapped)
            for (int s2 = 0; s2 < n; ++s2) norm0 = norm0 || mapped[s2] != 1.0;
        if (want_mean && dtype == DFB_I32 && !norm0 && filter == DFB_FILTER_MEAN && ctx->opt_filter_events && tiles > 0) {
            DFB_TRY(mean_tmp.alloc(ctx, (size_t)tiles * K1_TILE));
            P.mean_tmp = mean_tmp.p;
        }
    }
    if (tiles > 0) DFB_TRY(launch_filter(0));
    DFB_CUDA(cudaGetLastError());
    DFB_TRY(exclusive_scan_u32(ctx, tile_count.p, tile_off.p, tiles, total_dev.p, DFB_K_PREP));
    int64_t N = 0;
    DFB_TRY(d2h(ctx, &N, total_dev.p, sizeof N));

    // position runs of the new mask
    {
        DevBuf<uint32_t> rc;
        DevBuf<int64_t> ro, rt;
        DFB_TRY(rc.alloc(ctx, (size_t)std::max<int64_t>(W, 1)));
        DFB_TRY(ro.alloc(ctx, (size_t)std::max<int64_t>(W, 1)));
        DFB_TRY(rt.alloc(ctx, 1));
        if (W > 0) {
            KTimer t(ctx, DFB_K_PREP);
            mask_run_count_kernel<<<grid_for(W, 256, ctx->sm_count), 256, 0, ctx->stream>>>(mask.p, W, len, rc.p);
        }
        DFB_TRY(exclusive_scan_u32(ctx, rc.p, ro.p, W, rt.p, DFB_K_PREP));
        int64_t nr = 0;
        DFB_TRY(d2h(ctx, &nr, rt.p, sizeof nr));
        std::vector<int32_t> starts((size_t)nr);
        if (nr > 0) {
            DevBuf<int32_t> rs;
            DFB_TRY(rs.alloc(ctx, (size_t)nr));
            {
                KTimer t(ctx, DFB_K_PREP);
                mask_run_emit_kernel<<<grid_for(W, 256, ctx->sm_count), 256, 0, ctx->stream>>>(mask.p, W, len, ro.p,
                                                                                               rs.p);
            }
            DFB_TRY(d2h(ctx, starts.data(), rs.p, (size_t)nr * sizeof(int32_t)));
        }
        new_pos_len->clear();
        for (int64_t r = 0; r < nr; ++r) {
            const int64_t e = r + 1 < nr ? starts[(size_t)r + 1] : len;
            new_pos_len->push_back((int32_t)(e - starts[(size_t)r]));
        }
        uint32_t w0 = 0;
        if (W > 0) DFB_TRY(d2h(ctx, &w0, mask.p, sizeof w0));
        *new_pos_first = (int)(w0 & 1u);
    }

    dfb_cov* c = new dfb_cov();
    c->ctx = ctx;
    c->n = n;
    c->N = N;
    c->dtype = (mapped || dtype == DFB_F64) ? DFB_F64 : DFB_I32;
    int rc = alloc_matrix(c);
    if (rc >= 0 && want_mean) rc = c->mean.alloc(ctx, (size_t)std::max<int64_t>(N, 1));
    if (rc < 0) {
        delete c;
        return rc;
    }
    c->has_mean = want_mean;
    if (N > 0) {
        P.tile_off = tile_off.p;
        P.ld = c->ld;
        P.out = c->dtype == DFB_I32 ? (void*)c->xi.p : (void*)c->xd.p;
        P.out_is_double = c->dtype == DFB_F64;
        P.mean = want_mean ? c->mean.p : nullptr;
        rc = launch_filter(1);
        if (rc < 0) {
            delete c;
            return rc;
        }
    }
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);  // `up` buffers die with this frame
    if (e != cudaSuccess) {
        set_error("filter kernels: %s", cudaGetErrorString(e));
        delete c;
        return DFB_ECUDA;
    }
    *out = c;
    return DFB_OK;
}

// finalidx[index] <- newindex (R/filterData.R:159-161): scatter the new mask runs (over the kept
// positions of the previous index) into the previous index' TRUE runs.  Run-level bookkeeping.
static void compose_position(const int32_t* prev_len, int64_t prev_nruns, int prev_first,
                             const std::vector<int32_t>& new_len, int new_first, std::vector<int32_t>* out_len,
                             int* out_first) {
    std::vector<std::pair<int, int64_t>> runs;  // (value, length)
    auto push = [&](int v, int64_t l) {
        if (l <= 0) return;
        if (!runs.empty() && runs.back().first == v) runs.back().second += l;
        else runs.push_back({v, l});
    };
    size_t ni = 0;
    int64_t nrem = new_len.empty() ? 0 : new_len[0];
    int nv = new_first;
    int pv = prev_first ? 1 : 0;
    for (int64_t r = 0; r < prev_nruns; ++r, pv ^= 1) {
        int64_t l = prev_len[r];
        if (!pv) {
            push(0, l);
            continue;
        }
        while (l > 0 && ni < new_len.size()) {
            const int64_t take = std::min(l, nrem);
            push(nv, take);
            l -= take;
            nrem -= take;
            if (nrem == 0) {
                ++ni;
                nv ^= 1;
                nrem = ni < new_len.size() ? new_len[ni] : 0;
            }
        }
    }
    out_len->clear();
    *out_first = runs.empty() ? 0 : runs[0].first;
    for (auto& r : runs) {
        // split runs longer than INT32_MAX never occur (chromosome length is int32 in R)
        out_len->push_back((int32_t)r.second);
    }
}

// ------------------------------------------------------------------------------------------
// preprocess: means, group means, log2(x + scalefac)

template <typename V>
__global__ void __launch_bounds__(256) minmax_kernel(const V* __restrict__ x, int64_t ld, int64_t N, int n,
                                                     long long* __restrict__ mn, long long* __restrict__ mx) {
    long long lo = LLONG_MAX, hi = LLONG_MIN;
    const V* __restrict__ col = x + (size_t)blockIdx.y * ld;  // one sample per blockIdx.y: no index division
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        const long long v = (long long)col[i];
        lo = v < lo ? v : lo;
        hi = v > hi ? v : hi;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(mn, lo);
        atomicMax(mx, hi);
    }
}

struct PrepParams {
    const void* x;
    int64_t ld, N;
    int n;
    double scalefac;
    const double* lut;  // log2(k + scalefac) for k in [0, lut_n), host-built with libm
    int64_t lut_n;
    int lut_s;          // leading table entries staged in shared memory (<= lut_n)
    const int32_t* group;  // [n] or nullptr
    const double* group_size;  // [G]
    int G;
    double* y;      // [n][ld], or nullptr when the log2 matrix is not materialised (STORE_Y = false)
    double* mean;   // [N] or nullptr (already present)
    double* gmean;  // [G][ld]
    double* ymean;  // [N] (sum_s y[s]) / n in sample order (the F-stat contract's row mean), or nullptr
    int* out_of_range;  // set when an integer value falls outside the table (the caller re-runs)
};

// One thread per base, eight samples' loads in flight; the head of the log2 table sits in shared memory
// (coverage values are small almost everywhere: a table gather from L1 per element would bound the kernel).
// GR: group sums in registers for up to GR groups (0: shared-memory accumulators, any G).
template <typename V, bool STORE_Y, int GR>
__global__ void __launch_bounds__(256) preprocess_kernel(const PrepParams P) {
    extern __shared__ double prep_s[];  // [lut_s] table head, [G][blockDim] group sums (GR == 0), group ids [n]
    double* lut_s = prep_s;
    double* gs_s = prep_s + P.lut_s;
    unsigned char* grp_s = reinterpret_cast<unsigned char*>(gs_s + (GR == 0 ? (size_t)P.G * blockDim.x : 0));
    for (int e = threadIdx.x; e < P.lut_s; e += blockDim.x) lut_s[e] = P.lut[e];
    if (P.G > 0)
        for (int e = threadIdx.x; e < P.n; e += blockDim.x) grp_s[e] = (unsigned char)P.group[e];
    __syncthreads();
    const V* __restrict__ x = reinterpret_cast<const V*>(P.x);
    constexpr int LB = 8;
    const int n = P.n;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P.N; i += (int64_t)gridDim.x * blockDim.x) {
        double sum = 0.0, ysum = 0.0;
        double gr[GR > 0 ? GR : 1];
#pragma unroll
        for (int g = 0; g < (GR > 0 ? GR : 1); ++g) gr[g] = 0.0;
        if (GR == 0)
            for (int g = 0; g < P.G; ++g) gs_s[g * blockDim.x + threadIdx.x] = 0.0;
        for (int s0 = 0; s0 < n; s0 += LB) {
            V rawv[LB];
            // clamped row index instead of a predicate per load: the loads stay back to back
#pragma unroll
            for (int e = 0; e < LB; ++e) rawv[e] = __ldg(x + (size_t)min(s0 + e, n - 1) * P.ld + i);
            asm volatile("" ::: "memory");
#pragma unroll
            for (int e = 0; e < LB; ++e) {
                const int s = s0 + e;
                if (s < n) {
                    const V raw = rawv[e];
                    const double v = (double)raw;
                    sum = s == 0 ? v : __dadd_rn(sum, v);
                    if (GR > 0) {
                        if (P.G > 0) {
                            const int g = grp_s[s];
#pragma unroll
                            for (int gg = 0; gg < GR; ++gg)
                                if (g == gg) gr[gg] = __dadd_rn(gr[gg], v);
                        }
                    } else if (P.G > 0) {
                        double* a = &gs_s[grp_s[s] * blockDim.x + threadIdx.x];
                        *a = __dadd_rn(*a, v);
                    }
                    double yv;
                    if (sizeof(V) == 4 && P.lut) {
                        const int64_t k = (int64_t)raw;
                        if (k >= 0 && k < P.lut_n) {
                            yv = lut_s[min((int)k, P.lut_s - 1)];
                            if (k >= P.lut_s) yv = __ldg(P.lut + k);
                        } else {
                            yv = 0.0;
                            *P.out_of_range = 1;
                        }
                    } else {
                        yv = log2(__dadd_rn(v, P.scalefac));
                    }
                    ysum = s == 0 ? yv : __dadd_rn(ysum, yv);
                    if (STORE_Y) P.y[(size_t)s * P.ld + i] = yv;
                }
            }
        }
        if (P.mean) P.mean[i] = __ddiv_rn(sum, (double)n);
        if (P.ymean) P.ymean[i] = __ddiv_rn(ysum, (double)n);
        for (int g = 0; g < P.G; ++g) {
            double gsum = GR > 0 ? 0.0 : gs_s[g * blockDim.x + threadIdx.x];
#pragma unroll
            for (int gg = 0; gg < GR; ++gg)
                if (g == gg) gsum = gr[gg];
            P.gmean[(size_t)g * P.ld + i] = __ddiv_rn(gsum, P.group_size[g]);
        }
    }
}

// Integer coverage with a log2 table and at most two groups -- the hot configuration.  Sums are integer (exact, so
// identical to the left-to-right double sums while every partial sum stays below 2^53: n <= 255 values of 31 bits),
// the second group's sum is selected by a bit mask of the samples and the first one is the difference; one range /
// table-head test per eight samples instead of one per value.
template <bool STORE_Y>
__global__ void __launch_bounds__(256) preprocess_i32_kernel(const PrepParams P) {
    extern __shared__ double prep_s[];  // [lut_s] table head, then one byte of group-2 membership bits per 8 samples
    double* lut_s = prep_s;
    unsigned char* gm_s = reinterpret_cast<unsigned char*>(prep_s + P.lut_s);
    const int n = P.n;
    for (int e = threadIdx.x; e < P.lut_s; e += blockDim.x) lut_s[e] = P.lut[e];
    for (int e = threadIdx.x; e < (n + 7) / 8; e += blockDim.x) {
        unsigned m = 0;
        for (int b = 0; b < 8; ++b)
            if (P.G > 1 && e * 8 + b < n && P.group[e * 8 + b] == 1) m |= 1u << b;
        gm_s[e] = (unsigned char)m;
    }
    __syncthreads();
    const int32_t* __restrict__ x = reinterpret_cast<const int32_t*>(P.x);
    const int lut_n = (int)min(P.lut_n, (int64_t)INT32_MAX), lut_h = P.lut_s;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P.N; i += (int64_t)gridDim.x * blockDim.x) {
        long long isum = 0, g1 = 0;
        double ysum = 0.0;  // 0.0 + y == y bit for bit (log2 never returns -0.0)
        bool bad = false;
        auto step = [&](const int s0, auto full_tag) {
            constexpr bool FULL = decltype(full_tag)::value;
            int32_t raw[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) raw[e] = __ldg(x + (size_t)(FULL ? s0 + e : min(s0 + e, n - 1)) * P.ld + i);
            asm volatile("" ::: "memory");
            const unsigned gm = gm_s[s0 >> 3];
            int mx = raw[0], mn = raw[0];
#pragma unroll
            for (int e = 1; e < 8; ++e) {
                mx = max(mx, raw[e]);
                mn = min(mn, raw[e]);
            }
            double yv[8];
            if (mn >= 0 && mx < lut_h) {
#pragma unroll
                for (int e = 0; e < 8; ++e) yv[e] = lut_s[raw[e]];
            } else {  // rare: beyond the table head, or outside the table (flagged: the caller re-runs)
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int k = raw[e];
                    yv[e] = 0.0;
                    if (k >= 0 && k < lut_n) yv[e] = k < lut_h ? lut_s[k] : __ldg(P.lut + k);
                    else bad = true;
                }
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                if (FULL || s0 + e < n) {
                    isum += raw[e];
                    g1 += ((gm >> e) & 1u) ? raw[e] : 0;
                    ysum = __dadd_rn(ysum, yv[e]);
                    if (STORE_Y) P.y[(size_t)(s0 + e) * P.ld + i] = yv[e];
                }
            }
        };
        int s0 = 0;
        for (; s0 + 8 <= n; s0 += 8) step(s0, std::true_type());
        if (s0 < n) step(s0, std::false_type());
        if (P.mean) P.mean[i] = __ddiv_rn((double)isum, (double)n);
        if (P.ymean) P.ymean[i] = __ddiv_rn(ysum, (double)n);
        if (P.G > 0) P.gmean[i] = __ddiv_rn((double)(isum - g1), P.group_size[0]);
        if (P.G > 1) P.gmean[(size_t)P.ld + i] = __ddiv_rn((double)g1, P.group_size[1]);
        if (bad) *P.out_of_range = 1;
    }
}

static int launch_preprocess(dfb_ctx* ctx, const dfb_cov* c, PrepParams P, bool store_y) {
    P.lut_s = (c->dtype == DFB_I32 && P.lut) ? (int)std::min<int64_t>(LUT_SMEM, P.lut_n) : 0;
    const bool greg = P.G <= 2;  // group sums in registers
    const size_t smem = ((size_t)P.lut_s + (greg ? 0 : (size_t)P.G * 256)) * sizeof(double) + (size_t)round_up(c->n, 8);
    auto go = [&](auto kern) -> int {
        if (smem > 48 * 1024) DFB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int occ = 0;  // one resident wave, grid-stride loop inside (every block loads the table head once)
        DFB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, smem));
        const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(c->N, 256), (int64_t)ctx->sm_count * std::max(occ, 1)));
        KTimer t(ctx, DFB_K_PREP);
        kern<<<grid, 256, smem, ctx->stream>>>(P);
        return DFB_OK;
    };
    if (c->dtype == DFB_I32 && P.lut && P.G <= 2) {
        if (store_y) DFB_TRY(go(preprocess_i32_kernel<true>));
        else DFB_TRY(go(preprocess_i32_kernel<false>));
    } else if (c->dtype == DFB_I32) {
        if (store_y) DFB_TRY(greg ? go(preprocess_kernel<int32_t, true, 2>) : go(preprocess_kernel<int32_t, true, 0>));
        else DFB_TRY(greg ? go(preprocess_kernel<int32_t, false, 2>) : go(preprocess_kernel<int32_t, false, 0>));
    } else {
        DFB_TRY(greg ? go(preprocess_kernel<double, true, 2>) : go(preprocess_kernel<double, true, 0>));
    }
    DFB_CUDA(cudaGetLastError());
    return DFB_OK;
}

int ensure_lut(dfb_ctx* ctx, double scalefac) {
    constexpr int64_t CTX_LUT = 65536;
    if (ctx->lut_dev && ctx->lut_n == CTX_LUT && ctx->lut_scalefac == scalefac) return DFB_OK;
    if (!ctx->lut_dev) DFB_CUDA(cudaMalloc((void**)&ctx->lut_dev, CTX_LUT * sizeof(double)));
    ctx->lut_n = 0;
    // glibc log2, the function R calls: coverageProcessed is bit-identical to log2(x + scalefac) evaluated
    // by R on this host
    std::vector<double> h((size_t)CTX_LUT);
    for (int64_t k = 0; k < CTX_LUT; ++k) h[(size_t)k] = log2((double)k + scalefac);
    DFB_TRY(h2d(ctx, ctx->lut_dev, h.data(), (size_t)CTX_LUT * sizeof(double)));
    DFB_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->lut_n = CTX_LUT;
    ctx->lut_scalefac = scalefac;
    return DFB_OK;
}

// Materialise log2(x + scalefac) of a handle whose dfb_preprocess skipped it (integer coverage, every value
// inside the context table -- checked there).
int ensure_y(dfb_ctx* ctx, const dfb_cov* cov) {
    dfb_cov* c = const_cast<dfb_cov*>(cov);
    if (c->y.p || !c->y_lazy || c->N == 0) return DFB_OK;
    DFB_TRY(ensure_lut(ctx, c->scalefac));
    DFB_TRY(c->y.alloc(ctx, (size_t)c->n * c->ld));
    PrepParams P;
    memset(&P, 0, sizeof P);
    P.x = c->xi.p;
    P.ld = c->ld;
    P.N = c->N;
    P.n = c->n;
    P.scalefac = c->scalefac;
    P.lut = ctx->lut_dev;
    P.lut_n = ctx->lut_n;
    P.y = c->y.p;
    DevBuf<int> oor;  // cannot fire: dfb_preprocess saw every value inside the table
    DFB_TRY(oor.alloc(ctx, 1));
    DFB_TRY(oor.zero());
    P.out_of_range = oor.p;
    return launch_preprocess(ctx, c, P, true);
}

}  // namespace dfb

using namespace dfb;

extern "C" {

int dfb_filter_rle(dfb_ctx* ctx, int n, int dtype, const void* const* values, const int32_t* const* lengths,
                   const int64_t* nruns, int64_t len, const double* mapped_per_xm, int filter, double cutoff,
                   const int32_t* prev_len, int64_t prev_nruns, int prev_first, dfb_cov** out) {
    DFB_REQUIRE(ctx && out, "dfb_filter_rle: NULL argument");
    *out = nullptr;
    DFB_REQUIRE(filter == DFB_FILTER_NONE || filter == DFB_FILTER_ONE || filter == DFB_FILTER_MEAN,
                "filter %%in%% c('one', 'mean') is not TRUE");
    DFB_REQUIRE(len >= 0 && len <= INT32_MAX, "coverage length %lld out of range", (long long)len);
    if (prev_len) {
        int64_t kept = 0;
        int v = prev_first ? 1 : 0;
        for (int64_t r = 0; r < prev_nruns; ++r, v ^= 1)
            if (v) kept += prev_len[r];
        DFB_REQUIRE(kept == len, "sum(index) = %lld does not match the coverage length %lld", (long long)kept,
                    (long long)len);
    }
    std::vector<int32_t> npl;
    int npf = 0;
    dfb_cov* c = nullptr;
    DFB_TRY(filter_rle_impl(ctx, n, dtype, values, lengths, nruns, len, mapped_per_xm, filter, cutoff, true, &npl, &npf,
                            &c));
    int rc;
    if (prev_len) {
        std::vector<int32_t> fl;
        int ff = 0;
        compose_position(prev_len, prev_nruns, prev_first, npl, npf, &fl, &ff);
        rc = c->pos.build(ctx, fl.data(), (int64_t)fl.size(), ff);
    } else {
        rc = c->pos.build(ctx, npl.data(), (int64_t)npl.size(), npf);
    }
    if (rc < 0) {
        delete c;
        return rc;
    }
    if (c->pos.N != c->N) {
        set_error("internal error: position index has %lld TRUE positions but %lld rows were kept",
                  (long long)c->pos.N, (long long)c->N);
        delete c;
        return DFB_ECUDA;
    }
    *out = c;
    return DFB_OK;
}

static int cov_from_dense_impl(dfb_ctx* ctx, int n, int64_t N, int dtype, const void* cov, int64_t ld, bool on_device,
                               const int32_t* pos_len, int64_t pos_nruns, int pos_first, dfb_cov** out,
                               bool view = false) {
    DFB_REQUIRE(ctx && out && (cov || N == 0), "dfb_cov_from_dense: NULL argument");
    *out = nullptr;
    DFB_REQUIRE(n > 0 && N >= 0 && ld >= N, "dfb_cov_from_dense: bad shape (n=%d, N=%lld, ld=%lld)", n, (long long)N,
                (long long)ld);
    DFB_REQUIRE(dtype == DFB_I32 || dtype == DFB_F64, "unknown dtype %d", dtype);
    DFB_REQUIRE(pos_len != nullptr, "'coverageInfo$position' is NULL when it shouldn't be.");
    dfb_cov* c = new dfb_cov();
    c->ctx = ctx;
    c->n = n;
    c->N = N;
    c->dtype = dtype;
    // the matrix copy is queued first: the device moves the coverage while the host builds the position
    // index below (its run loops and small uploads used to leave the GPU idle at the start of a step)
    // a device matrix in the engine's own layout (row stride a multiple of 64 elements, 256-byte aligned
    // base) is read in place when the caller asks for a view: no second copy of resident coverage
    view = view && on_device && N > 0 && ld % 64 == 0 && ((uintptr_t)cov & 255u) == 0;
    int rc = DFB_OK;
    if (view) {
        c->ld = ld;
        if (dtype == DFB_I32) c->xi.borrow(ctx->stream, (int32_t*)const_cast<void*>(cov), (size_t)n * ld);
        else c->xd.borrow(ctx->stream, (double*)const_cast<void*>(cov), (size_t)n * ld);
    } else {
        rc = alloc_matrix(c);
    }
    if (rc < 0) {
        delete c;
        return rc;
    }
    const size_t es = dtype == DFB_I32 ? 4 : 8;
    void* dst = dtype == DFB_I32 ? (void*)c->xi.p : (void*)c->xd.p;
    if (N > 0 && !view) {
        cudaError_t e = cudaMemcpy2DAsync(dst, (size_t)c->ld * es, cov, (size_t)ld * es, (size_t)N * es, (size_t)n,
                                          on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ctx->stream);
        if (e != cudaSuccess) {
            set_error("coverage upload: %s", cudaGetErrorString(e));
            delete c;
            return DFB_ECUDA;
        }
    }
    rc = c->pos.build(ctx, pos_len, pos_nruns, pos_first);
    if (rc >= 0 && c->pos.N != N) {
        set_error("nrow(coverage) = %lld does not match sum(position) = %lld", (long long)N, (long long)c->pos.N);
        rc = DFB_EINVAL;
    }
    if (rc < 0) {
        cudaStreamSynchronize(ctx->stream);
        delete c;
        return rc;
    }
    *out = c;
    return DFB_OK;
}

int dfb_cov_from_dense(dfb_ctx* ctx, int n, int64_t N, int dtype, const void* cov, int64_t ld, const int32_t* pos_len,
                       int64_t pos_nruns, int pos_first, dfb_cov** out) {
    int rc = cov_from_dense_impl(ctx, n, N, dtype, cov, ld, false, pos_len, pos_nruns, pos_first, out);
    if (rc >= 0) cudaStreamSynchronize(ctx->stream);  // the caller may free `cov` right away
    return rc;
}

int dfb_cov_from_dense_dev(dfb_ctx* ctx, int n, int64_t N, int dtype, const void* cov_dev, int64_t ld,
                           const int32_t* pos_len, int64_t pos_nruns, int pos_first, dfb_cov** out) {
    return cov_from_dense_impl(ctx, n, N, dtype, cov_dev, ld, true, pos_len, pos_nruns, pos_first, out);
}

int dfb_cov_view_dense_dev(dfb_ctx* ctx, int n, int64_t N, int dtype, const void* cov_dev, int64_t ld,
                           const int32_t* pos_len, int64_t pos_nruns, int pos_first, dfb_cov** out) {
    return cov_from_dense_impl(ctx, n, N, dtype, cov_dev, ld, true, pos_len, pos_nruns, pos_first, out, true);
}

int dfb_cov_from_rle(dfb_ctx* ctx, int n, int dtype, const void* const* values, const int32_t* const* lengths,
                     const int64_t* nruns, int64_t N, const int32_t* pos_len, int64_t pos_nruns, int pos_first,
                     dfb_cov** out) {
    DFB_REQUIRE(ctx && out, "dfb_cov_from_rle: NULL argument");
    *out = nullptr;
    DFB_REQUIRE(pos_len != nullptr, "'coverageInfo$position' is NULL when it shouldn't be.");
    dfb_cov* c = new dfb_cov();
    c->ctx = ctx;
    c->n = n;
    c->N = N;
    c->dtype = dtype;
    int rc = alloc_matrix(c);
    if (rc < 0) {
        delete c;
        return rc;
    }
    // The samples travel in groups: upload (host pool packs / narrows into the pinned ring, DMA), prefix sums,
    // validation, run table and expansion of a group are queued without a host synchronisation, so the device
    // expands group g while the host is still packing group g + 1 (the upload is bound by the host pass).
    int64_t total_runs = 0;
    for (int s = 0; s < n; ++s) total_runs += nruns[s];
    const int G = (N > 0 && n >= 8 && total_runs >= (1 << 20)) ? 4 : 1;
    struct Group {
        RleUpload up;
        DevBuf<int64_t> tab;  // run holding the first position of every (tile, sample)
        int s0 = 0, ns = 0;
    };
    std::vector<std::unique_ptr<Group>> groups;
    DevBuf<int64_t> bad;
    std::vector<cudaEvent_t> ev_dma;  // page-locked path: one event per group's DMAs
    auto drop_events = [&]() {
        for (auto e : ev_dma) cudaEventDestroy(e);
        ev_dma.clear();
    };
    auto fail = [&](int code) -> int {
        if (ctx->aux) cudaStreamSynchronize(ctx->aux);  // DMAs of the page-locked path
        cudaStreamSynchronize(ctx->stream);              // queued kernels read the groups' buffers
        drop_events();
        delete c;
        return code;
    };
    rc = bad.alloc(ctx, (size_t)2 * G);
    if (rc < 0) return fail(rc);
    if (cudaMemsetAsync(bad.p, 0xff, (size_t)2 * G * sizeof(int64_t), ctx->stream) != cudaSuccess) return fail(DFB_ECUDA);
    // 1024-position tiles on 128 threads, 16 blocks per SM
    constexpr int XT = 128, XTILE = XT * K1_PER;
    const int64_t tiles = ceil_div(N, XTILE);
    // Page-locked caller memory (cudaHostAlloc / cudaHostRegister): the columns are DMAed where they lie, on the
    // auxiliary stream so that group g + 1 travels while group g is expanded -- no host core touches the data, which
    // matters most when eight ranks share the host.  Pageable memory goes through the pinned ring (above).
    bool pinned = N > 0 && total_runs >= (1 << 16) && ctx->opt_pinned_direct && ensure_aux_stream(ctx) >= 0;
    for (int s = 0; s < n && pinned; ++s) pinned = nruns[s] == 0 || (host_is_pinned(values[s]) && host_is_pinned(lengths[s]));
    std::vector<int64_t*> offs((size_t)G, nullptr);
    for (int g = 0; g < G; ++g) {
        groups.emplace_back(new Group());
        Group& gr = *groups.back();
        gr.s0 = (int)((int64_t)n * g / G);
        gr.ns = (int)((int64_t)n * (g + 1) / G) - gr.s0;
        if (pinned) {  // every buffer first: the DMAs on the other stream must not wait behind this stream's kernels
            rc = rle_alloc(ctx, gr.ns, dtype, values + gr.s0, lengths + gr.s0, nruns + gr.s0, N, &gr.up, &offs[(size_t)g]);
            if (rc < 0) return fail(rc);
        }
    }
    if (pinned) {
        bool ok = cudaEventRecord(ctx->ev_main, ctx->stream) == cudaSuccess &&
                  cudaStreamWaitEvent(ctx->aux, ctx->ev_main, 0) == cudaSuccess;
        for (int g = 0; g < G && ok; ++g) {
            Group& gr = *groups[(size_t)g];
            cudaEvent_t e = nullptr;
            ok = cudaEventCreateWithFlags(&e, cudaEventDisableTiming) == cudaSuccess;
            if (!ok) break;
            ev_dma.push_back(e);
            ok = rle_transfer_pinned(gr.ns, dtype, values + gr.s0, lengths + gr.s0, nruns + gr.s0, offs[(size_t)g], &gr.up,
                                     ctx->aux) >= 0 &&
                 cudaEventRecord(e, ctx->aux) == cudaSuccess;
        }
        if (!ok) {
            set_error("Rle upload from page-locked memory failed: %s", cudaGetErrorString(cudaGetLastError()));
            return fail(DFB_ECUDA);
        }
    }
    for (int g = 0; g < G; ++g) {
        Group& gr = *groups[(size_t)g];
        if (pinned) {
            rc = cudaStreamWaitEvent(ctx->stream, ev_dma[(size_t)g], 0) == cudaSuccess ? DFB_OK : DFB_ECUDA;
            if (rc >= 0) rc = rle_finish(ctx, gr.ns, N, offs[(size_t)g], &gr.up, bad.p + 2 * g);
        } else {
            rc = upload_rle_async(ctx, gr.ns, dtype, values + gr.s0, lengths + gr.s0, nruns + gr.s0, N, &gr.up, /*try16=*/true,
                                  bad.p + 2 * g);
        }
        if (rc < 0) return fail(rc);
        if (N == 0) continue;
        RleDev R;
        R.values = gr.up.values.p;
        R.cum = gr.up.cum.p;
        R.run_off = gr.up.run_off.p;
        R.n = gr.ns;
        R.len = N;
        rc = gr.tab.alloc(ctx, (size_t)tiles * gr.ns);
        if (rc < 0) return fail(rc);
        const dim3 grid((unsigned)std::min<int64_t>(tiles, (int64_t)ctx->sm_count * 16), (unsigned)gr.ns);
        KTimer t(ctx, DFB_K_PREP, 2);
        tile_runs_kernel<<<grid_for(tiles * gr.ns, 256, ctx->sm_count, 16), 256, 0, ctx->stream>>>(R, tiles, gr.tab.p, XTILE);
        if (dtype == DFB_I32)
            expand_rle_tiles_kernel<int32_t, XT><<<grid, XT, 0, ctx->stream>>>(R, tiles, gr.tab.p, c->ld,
                                                                               c->xi.p + (size_t)gr.s0 * c->ld, bad.p + 2 * g);
        else
            expand_rle_tiles_kernel<double, XT><<<grid, XT, 0, ctx->stream>>>(R, tiles, gr.tab.p, c->ld,
                                                                              c->xd.p + (size_t)gr.s0 * c->ld, bad.p + 2 * g);
    }
    if (cudaGetLastError() != cudaSuccess) {
        set_error("Rle expansion: launch failed");
        return fail(DFB_ECUDA);
    }
    // the position index is built while the expansion runs (host run loops + two small uploads)
    rc = c->pos.build(ctx, pos_len, pos_nruns, pos_first);
    if (rc >= 0 && c->pos.N != N) {
        set_error("nrow(coverage) = %lld does not match sum(position) = %lld", (long long)N, (long long)c->pos.N);
        rc = DFB_EINVAL;
    }
    if (rc < 0) return fail(rc);
    std::vector<int64_t> hbad((size_t)2 * G, -1);
    rc = d2h(ctx, hbad.data(), bad.p, hbad.size() * sizeof(int64_t));  // synchronises: the groups' buffers die with this frame
    drop_events();
    if (rc < 0) return fail(rc);
    for (int g = 0; g < G; ++g) {
        rc = check_rle_flags(hbad.data() + 2 * g, groups[(size_t)g]->s0, N);
        if (rc < 0) return fail(rc);
    }
    *out = c;
    return DFB_OK;
}

void dfb_cov_free(dfb_cov* cov) {
    if (!cov) return;
    bg_download_wait(cov->ctx);  // a pending dfb_fstats_begin download reads this handle's buffer
    cudaStreamSynchronize(cov->ctx->stream);
    delete cov;
}

int dfb_cov_info(const dfb_cov* c, int64_t* N, int* n, int64_t* chr_len, int* dtype, int64_t* pos_nruns,
                 int* n_groups, int* has_processed) {
    DFB_REQUIRE(c, "dfb_cov_info: NULL handle");
    if (N) *N = c->N;
    if (n) *n = c->n;
    if (chr_len) *chr_len = c->pos.chr_len;
    if (dtype) *dtype = c->dtype;
    if (pos_nruns) *pos_nruns = (int64_t)c->pos.len.size();
    if (n_groups) *n_groups = c->G;
    if (has_processed) *has_processed = c->y.p != nullptr || c->y_lazy;
    return DFB_OK;
}

int dfb_cov_get_position(const dfb_cov* c, int32_t* lengths, int* first) {
    DFB_REQUIRE(c, "dfb_cov_get_position: NULL handle");
    if (lengths) memcpy(lengths, c->pos.len.data(), c->pos.len.size() * sizeof(int32_t));
    if (first) *first = c->pos.first;
    return DFB_OK;
}

int dfb_cov_get_column(const dfb_cov* c, int sample, void* out) {
    DFB_REQUIRE(c && out, "dfb_cov_get_column: NULL argument");
    DFB_REQUIRE(sample >= 0 && sample < c->n, "sample %d out of range", sample);
    if (c->dtype == DFB_I32) return d2h_big(c->ctx, out, c->xi.p + (size_t)sample * c->ld, (size_t)c->N * 4);
    return d2h_big(c->ctx, out, c->xd.p + (size_t)sample * c->ld, (size_t)c->N * 8);
}

int dfb_cov_get_mean(const dfb_cov* c, double* out) {
    DFB_REQUIRE(c && out, "dfb_cov_get_mean: NULL argument");
    if (!c->has_mean) {
        set_error("meanCoverage is not available (filter with returnMean or run dfb_preprocess)");
        return DFB_ESTATE;
    }
    return d2h_big(c->ctx, out, c->mean.p, (size_t)c->N * sizeof(double));
}

int dfb_preprocess(dfb_ctx* ctx, dfb_cov* c, double scalefac, const int32_t* group_of_sample, int n_groups) {
    DFB_REQUIRE(ctx && c, "dfb_preprocess: NULL argument");
    DFB_REQUIRE(n_groups >= 0 && n_groups <= 64, "at most 64 groups are supported (got %d)", n_groups);
    DFB_REQUIRE(n_groups == 0 || group_of_sample, "groupInfo is NULL but n_groups > 0");
    std::vector<double> gsize((size_t)n_groups, 0.0);
    for (int s = 0; s < c->n && n_groups > 0; ++s) {
        DFB_REQUIRE(group_of_sample[s] >= 0 && group_of_sample[s] < n_groups, "groupInfo[%d] out of range", s + 1);
        gsize[(size_t)group_of_sample[s]] += 1.0;
    }
    c->scalefac = scalefac;
    c->G = n_groups;
    // integer coverage: the log2 matrix is not written (dfb_cov::y_lazy) unless a value falls outside the table
    // (only when the consumer will be the second-generation null scan: the first-generation kernels read the matrix)
    bool lazy = ctx->opt_lazy_y && c->dtype == DFB_I32 && null2_wanted(ctx, c->n);
    c->y.release();
    c->y_lazy = false;
    c->has_ym = false;
    if (!lazy) DFB_TRY(c->y.alloc(ctx, (size_t)c->n * c->ld));
    DFB_TRY(c->ym.alloc(ctx, (size_t)std::max<int64_t>(c->N, 1)));
    if (n_groups > 0) DFB_TRY(c->gmean.alloc(ctx, (size_t)n_groups * c->ld));
    const bool need_mean = !c->has_mean;
    if (need_mean) DFB_TRY(c->mean.alloc(ctx, (size_t)std::max<int64_t>(c->N, 1)));
    if (c->N == 0) {
        c->has_mean = true;
        c->has_ym = true;
        c->y_lazy = lazy;
        return DFB_OK;
    }
    // integer coverage: log2 through a host-built table (glibc log2, the function R calls), so
    // coverageProcessed is bit-identical to log2(x + scalefac) evaluated by R on this host.  The
    // context keeps a 64 Ki-entry table per scalefac; a value outside it (flagged by the kernel) falls
    // back to a table sized from the data's min/max, or to device log2 beyond 2^22.
    DevBuf<double> lut;
    DevBuf<int> oor;
    DFB_TRY(oor.alloc(ctx, 1));
    DFB_TRY(oor.zero());
    if (c->dtype == DFB_I32) DFB_TRY(ensure_lut(ctx, scalefac));
    DevBuf<int32_t> grp;
    DevBuf<double> gsz;
    if (n_groups > 0) {
        DFB_TRY(grp.alloc(ctx, c->n));
        DFB_TRY(gsz.alloc(ctx, n_groups));
        DFB_TRY(h2d(ctx, grp.p, group_of_sample, (size_t)c->n * sizeof(int32_t)));
        DFB_TRY(h2d(ctx, gsz.p, gsize.data(), (size_t)n_groups * sizeof(double)));
    }
    PrepParams P;
    memset(&P, 0, sizeof P);
    P.x = c->dtype == DFB_I32 ? (const void*)c->xi.p : (const void*)c->xd.p;
    P.ld = c->ld;
    P.N = c->N;
    P.n = c->n;
    P.scalefac = scalefac;
    P.lut = c->dtype == DFB_I32 ? ctx->lut_dev : nullptr;
    P.lut_n = c->dtype == DFB_I32 ? ctx->lut_n : 0;
    P.group = n_groups ? grp.p : nullptr;
    P.group_size = n_groups ? gsz.p : nullptr;
    P.G = n_groups;
    P.y = c->y.p;
    P.mean = need_mean ? c->mean.p : nullptr;
    P.gmean = n_groups ? c->gmean.p : nullptr;
    P.ymean = c->ym.p;
    P.out_of_range = oor.p;
    DFB_TRY(launch_preprocess(ctx, c, P, !lazy));
    int flagged = 0;
    DFB_TRY(d2h(ctx, &flagged, oor.p, sizeof flagged));  // synchronises: group arrays are locals
    if (flagged) {
        // rare: coverage beyond the context table.  Size a table from the data, or use device log2; the log2
        // matrix is materialised in this case (the lazy readers only know the context table).
        DevBuf<long long> mm;
        DFB_TRY(mm.alloc(ctx, 2));
        long long init[2] = {LLONG_MAX, LLONG_MIN};
        DFB_TRY(h2d(ctx, mm.p, init, sizeof init));
        {
            KTimer t(ctx, DFB_K_PREP);
            minmax_kernel<int32_t><<<dim3(grid_for(c->N, 256 * 8, ctx->sm_count, 4), c->n), 256, 0, ctx->stream>>>(
                c->xi.p, c->ld, c->N, c->n, mm.p, mm.p + 1);
        }
        long long got[2];
        DFB_TRY(d2h(ctx, got, mm.p, sizeof got));
        P.lut = nullptr;
        P.lut_n = 0;
        if (got[0] >= 0 && got[1] < (1ll << 22)) {
            std::vector<double> h((size_t)got[1] + 1);
            for (int64_t k = 0; k <= got[1]; ++k) h[(size_t)k] = log2((double)k + scalefac);
            DFB_TRY(lut.alloc(ctx, h.size()));
            DFB_TRY(h2d(ctx, lut.p, h.data(), h.size() * sizeof(double)));
            DFB_CUDA(cudaStreamSynchronize(ctx->stream));
            P.lut = lut.p;
            P.lut_n = got[1] + 1;
        }
        if (lazy) {
            lazy = false;
            DFB_TRY(c->y.alloc(ctx, (size_t)c->n * c->ld));
            P.y = c->y.p;
        }
        DFB_TRY(oor.zero());
        DFB_TRY(launch_preprocess(ctx, c, P, true));
        DFB_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    c->has_mean = true;
    c->has_ym = true;
    c->y_lazy = lazy;
    c->has_f = false;
    c->a16_center = -1;  // the log2 matrix changed: the null scan's fp16 operand is stale
    return DFB_OK;
}

int dfb_cov_get_processed_column(const dfb_cov* c, int sample, double* out) {
    DFB_REQUIRE(c && out, "dfb_cov_get_processed_column: NULL argument");
    DFB_REQUIRE(sample >= 0 && sample < c->n, "sample %d out of range", sample);
    if (!c->y.p && !c->y_lazy) {
        set_error("coverageProcessed is not available: run dfb_preprocess first");
        return DFB_ESTATE;
    }
    DFB_TRY(ensure_y(c->ctx, c));
    return d2h_big(c->ctx, out, c->y.p + (size_t)sample * c->ld, (size_t)c->N * sizeof(double));
}

int dfb_cov_get_group_mean(const dfb_cov* c, int group, double* out) {
    DFB_REQUIRE(c && out, "dfb_cov_get_group_mean: NULL argument");
    DFB_REQUIRE(group >= 0 && group < c->G, "group %d out of range (%d groups)", group, c->G);
    return d2h_big(c->ctx, out, c->gmean.p + (size_t)group * c->ld, (size_t)c->N * sizeof(double));
}

int64_t dfb_chunk_lengths(int64_t N, double chunksize, int mc_cores, int64_t* lens, int64_t cap) {
    // R/preprocessCoverage.R:171-188, 218-227
    if (!(chunksize > 0)) chunksize = ceil((double)N / (double)(mc_cores > 0 ? mc_cores : 1));
    int64_t lastloop = chunksize > 0 ? (int64_t)trunc((double)N / chunksize) : 0;
    if (chunksize > 0 && fmod((double)N, chunksize) == 0 && lastloop > 0) lastloop -= 1;
    if (lastloop == 0) {
        if (lens && cap > 0) lens[0] = N;
        return 1;
    }
    int64_t used = 0, k = 0;
    for (; k < lastloop; ++k) {
        if (lens && k < cap) lens[k] = (int64_t)chunksize;
        used += (int64_t)chunksize;
    }
    if (N - used > 0) {
        if (lens && k < cap) lens[k] = N - used;
        ++k;
    }
    return k;
}

}  // extern "C"
