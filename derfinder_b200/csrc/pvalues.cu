// K5 and the permutation driver: null regions per permutation, null-area sort, exact p-value and
// FWER counting, empirical quantile cutoff.
//
// Reference: calculatePvalues (R/calculatePvalues.R:157-430; the permutation loop :306-358, the
// duplicated null storage :346-354, the stable decreasing-area order :368), .calcPval (:432-441),
// .calculateFWER (R/mergeResults.R:380-386), quantile(fstats, 0.99) default cutoff
// (R/findRegions.R:118).  Sorting uses cub::DeviceRadixSort (CUDA toolkit library).
#include <math.h>

#include <algorithm>
#include <functional>
#include <cub/cub.cuh>

#include "data.cuh"

namespace dfb {

// ------------------------------------------------------------------------------------------

__global__ void add_perm_offset_kernel(int32_t* __restrict__ perm, int64_t n, int32_t offset) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) perm[i] += offset;
}

static int perm_nulls_impl(dfb_ctx* ctx, dfb_cov* cov, const double* mod, int p, const double* mod0, int p0,
                           double adjust_f, const int32_t* perm_idx, int64_t B, double lo, double hi,
                           int32_t max_region_gap, dfb_nulls** out,
                           const std::function<void()>* after_first_launch = nullptr) {
    const bool after_first_launch_given = after_first_launch != nullptr;
    *out = nullptr;
    DFB_REQUIRE(cov->y.p != nullptr || cov->y_lazy, "coverageProcessed is missing: run dfb_preprocess first");
    DFB_REQUIRE(B >= 0 && (perm_idx || B == 0), "perm_idx is NULL");
    DFB_REQUIRE(lo <= hi, "cutoff pair must be sorted (L <= U)");
    ModelBasis mb;
    DFB_TRY(build_model_basis(mod, cov->n, p, mod0, p0, &mb));
    const int n = cov->n;
    for (int64_t b = 0; b < B; ++b)
        for (int k = 0; k < n; ++k)
            DFB_REQUIRE(perm_idx[b * n + k] >= 0 && perm_idx[b * n + k] < n,
                        "perm_idx[%lld, %d] = %d is not a 0-based sample index", (long long)b, k, perm_idx[b * n + k]);
    dfb_nulls* nl = new dfb_nulls();
    nl->ctx = ctx;
    nl->B = B;
    nl->per_perm.assign((size_t)B, 0);
    if (B == 0 || cov->N == 0) {
        *out = nl;
        return DFB_OK;
    }
    struct Guard {
        dfb_nulls* p;
        ~Guard() { delete p; }
    } guard{nl};

    DevBuf<double> qd;
    DevBuf<int32_t> idx_d;
    DevBuf<uint32_t> seg;
    DFB_TRY(qd.alloc(ctx, mb.qpad.size()));
    DFB_TRY(h2d(ctx, qd.p, mb.qpad.data(), mb.qpad.size() * sizeof(double)));
    DFB_TRY(idx_d.alloc(ctx, (size_t)B * n));
    DFB_TRY(h2d(ctx, idx_d.p, perm_idx, (size_t)B * n * sizeof(int32_t)));
    DFB_TRY(build_segment_starts(ctx, cov->pos, max_region_gap, &seg));

    // tensor-core screen + exact refinement when the cutoff is one-sided positive (the normal case);
    // otherwise the dense FP64 kernel.  Results are identical either way.
    const bool use_null2 = null2_supported(ctx, cov, mb, lo, hi, adjust_f);
    const bool use_fused = use_null2 || fused_supported(ctx, cov, mb, lo, hi, adjust_f);
    const bool use_screen = use_fused || screen_supported(ctx, cov, mb, lo, hi, adjust_f);
    DevBuf<double> q_d;
    if (use_screen) {
        DFB_TRY(q_d.alloc(ctx, mb.q.size()));
        DFB_TRY(h2d(ctx, q_d.p, mb.q.data(), mb.q.size() * sizeof(double)));
    }
    // F >= 0 (or NaN) by construction, so the "dn" side (F <= L) can only fire when L >= 0
    const bool two_sided = lo >= 0.0;
    const int sides = two_sided ? 2 : 1;
    const int64_t N = cov->N;
    const int tb = use_screen ? 32 : fstat_tile_bases(n);  // bases per value-offset entry
    // per permutation: mask words, value offsets, and per 32-word group the region counts + offsets (+ flag)
    const double per_perm_fixed = (double)sides * ((double)ceil_div(N, 32) * 4.0 + (double)ceil_div(N, tb) * 8.0 +
                                                   (double)ceil_div(N, 1024) * 13.0);
    const double budget = ctx->opt_scratch_gb * 1e9;
    int64_t pb = (int64_t)std::max(1.0, floor(budget * 0.5 / per_perm_fixed));
    if (ctx->opt_perm_batch > 0) pb = ctx->opt_perm_batch;
    pb = std::min<int64_t>(pb, B);
    if (ctx->opt_perm_batch <= 0) pb = ceil_div(B, ceil_div(B, pb));  // batches of equal size

    std::vector<RegionSet> sets;
    std::vector<std::pair<int64_t, int64_t>> work;  // [b_lo, b_hi)
    for (int64_t b = B; b > 0;) {
        int64_t lo_b = std::max<int64_t>(0, b - pb);
        work.push_back({lo_b, b});
        b = lo_b;
    }
    // work is a stack: batches come off in increasing permutation order
    // exceedance density: the context remembers what the last scan saw (a fresh context assumes 8 %), so a
    // repeated analysis does not reserve gigabytes of value slots it will not use; an underestimate costs one
    // re-run of the scan with the exact capacity
    double density = ctx->opt_null_cap_mb > 0 ? -1.0 : (ctx->null_density > 0 ? ctx->null_density : 0.08);
    while (!work.empty()) {
        auto [b_lo, b_hi] = work.back();
        work.pop_back();
        const int64_t nb = b_hi - b_lo;
        const double dense = (double)nb * sides * (double)N;
        size_t hint = density < 0 ? (size_t)(ctx->opt_null_cap_mb * 1e6 / 8.0)
                                  : (size_t)std::max(dense * density, 4.0e6);
        if ((double)hint * 8.0 > budget * 0.5) hint = (size_t)(budget * 0.5 / 8.0);
        PermBatch pbatch;
        int rc = use_null2  ? fstat_perm_batch_null2(ctx, cov, mb, q_d.p, adjust_f, idx_d.p + (size_t)b_lo * n, nb, hi,
                                                     hint, &pbatch, true)
                 : use_fused  ? fstat_perm_batch_fused(ctx, cov, mb, q_d.p, adjust_f, idx_d.p + (size_t)b_lo * n, nb, hi,
                                                     hint, &pbatch, true)
                 : use_screen ? fstat_perm_batch_screen(ctx, cov, mb, q_d.p, adjust_f, idx_d.p + (size_t)b_lo * n, nb, hi,
                                                      hint, &pbatch, nullptr)
                            : fstat_perm_batch(ctx, cov, mb, qd.p, adjust_f, idx_d.p + (size_t)b_lo * n, nb, lo, hi,
                                               two_sided, hint, &pbatch);
        if (rc == DFB_ENOMEM && nb > 1) {  // exceedances did not fit: halve the batch
            const int64_t mid = b_lo + nb / 2;
            work.push_back({mid, b_hi});
            work.push_back({b_lo, mid});
            continue;
        }
        DFB_TRY(rc);
        if (after_first_launch) {  // the null scan is in flight: independent work can be enqueued behind it now
            (*after_first_launch)();
            after_first_launch = nullptr;
        }
        RegionSet rs;
        DFB_TRY(find_regions_rows(ctx, N, pbatch.rows, sides, pbatch.mask.p, seg.p, nullptr, 0, &pbatch, false, true,
                                  &rs));
        if (rs.total < 0) {
            // deferred overflow check of the fused path fired: re-run K2 with the exact capacity
            const size_t need = (size_t)*pbatch.used_host;
            pbatch = PermBatch();
            rc = use_null2 ? fstat_perm_batch_null2(ctx, cov, mb, q_d.p, adjust_f, idx_d.p + (size_t)b_lo * n, nb, hi, need,
                                                    &pbatch, false)
                           : fstat_perm_batch_fused(ctx, cov, mb, q_d.p, adjust_f, idx_d.p + (size_t)b_lo * n, nb, hi, need,
                                                    &pbatch, false);
            if (rc == DFB_ENOMEM && nb > 1) {
                const int64_t mid = b_lo + nb / 2;
                work.push_back({mid, b_hi});
                work.push_back({b_lo, mid});
                continue;
            }
            DFB_TRY(rc);
            DFB_TRY(find_regions_rows(ctx, N, pbatch.rows, sides, pbatch.mask.p, seg.p, nullptr, 0, &pbatch, false,
                                      true, &rs));
        }
        DFB_TRY(null2_release_mask(ctx, &pbatch));  // the mask rows have been consumed: hand the kept-zero buffer back clean
        if (pbatch.used_host) pbatch.nvals = (int64_t)*pbatch.used_host;
        if (density >= 0 && dense > 0) {
            density = std::min(1.0, std::max(1e-4, 1.3 * (double)pbatch.nvals / dense));
            ctx->null_density = density;
        }
        for (int64_t b = 0; b < nb; ++b)
            for (int s = 0; s < sides; ++s) nl->per_perm[(size_t)(b_lo + b)] += rs.row_counts[(size_t)(b * sides + s)];
        if (rs.total > 0 && b_lo > 0) {
            KTimer t(ctx, DFB_K_REGIONS);
            add_perm_offset_kernel<<<(unsigned)ceil_div(rs.total, 256), 256, 0, ctx->stream>>>(rs.perm.p, rs.total,
                                                                                              (int32_t)b_lo);
        }
        sets.push_back(std::move(rs));
    }
    int64_t total = 0;
    for (auto& s : sets) total += s.total;
    nl->count = total;
    if (total > 0) {
        if (sets.size() == 1) {
            nl->area = std::move(sets[0].area);
            nl->stat = std::move(sets[0].stat);
            nl->width = std::move(sets[0].width);
            nl->perm = std::move(sets[0].perm);
        } else {
            DFB_TRY(nl->area.alloc(ctx, (size_t)total));
            DFB_TRY(nl->stat.alloc(ctx, (size_t)total));
            DFB_TRY(nl->width.alloc(ctx, (size_t)total));
            DFB_TRY(nl->perm.alloc(ctx, (size_t)total));
            int64_t o = 0;
            for (auto& s : sets) {
                if (s.total == 0) continue;
                DFB_CUDA(cudaMemcpyAsync(nl->area.p + o, s.area.p, (size_t)s.total * 8, cudaMemcpyDeviceToDevice, ctx->stream));
                DFB_CUDA(cudaMemcpyAsync(nl->stat.p + o, s.stat.p, (size_t)s.total * 8, cudaMemcpyDeviceToDevice, ctx->stream));
                DFB_CUDA(cudaMemcpyAsync(nl->width.p + o, s.width.p, (size_t)s.total * 4, cudaMemcpyDeviceToDevice, ctx->stream));
                DFB_CUDA(cudaMemcpyAsync(nl->perm.p + o, s.perm.p, (size_t)s.total * 4, cudaMemcpyDeviceToDevice, ctx->stream));
                o += s.total;
            }
        }
    }
    // stand-alone callers get a finished result; dfb_calculate_pvalues keeps enqueuing behind it
    if (!after_first_launch_given) DFB_CUDA(cudaStreamSynchronize(ctx->stream));
    guard.p = nullptr;
    *out = nl;
    return DFB_OK;
}

// ------------------------------------------------------------------------------------------
// p-values: sorted nulls + upper bound

__global__ void pval_kernel(const double* __restrict__ areas, int64_t R, const double* __restrict__ sorted_nulls,
                            int64_t M, double weight, double* __restrict__ p) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const double a = areas[r];
    // k = number of nulls <= a  (findInterval on the unique values, then "strictly greater" count)
    int64_t lo = 0, hi = M;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (sorted_nulls[mid] <= a) lo = mid + 1;
        else hi = mid;
    }
    p[r] = __ddiv_rn(weight * (double)(M - lo) + 1.0, weight * (double)M + 1.0);
}

static int sort_doubles(dfb_ctx* ctx, const double* in, double* out, int64_t n) {
    size_t tmp_bytes = 0;
    DFB_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, in, out, n, 0, 64, ctx->stream));
    DevBuf<unsigned char> tmp;
    DFB_TRY(tmp.alloc(ctx, tmp_bytes));
    {
        KTimer t(ctx, DFB_K_PVAL, 8);
        DFB_CUDA(cub::DeviceRadixSort::SortKeys(tmp.p, tmp_bytes, in, out, n, 0, 64, ctx->stream));
    }
    return DFB_OK;
}

// number of sorted nulls strictly greater than each area, ADDED to counts (multi-GPU pooling: one
// call per rank's sorted null list)
__global__ void count_greater_kernel(const double* __restrict__ areas, int64_t R, const double* __restrict__ sorted_nulls,
                                     int64_t M, int64_t* __restrict__ counts) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const double a = areas[r];
    int64_t lo = 0, hi = M;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (sorted_nulls[mid] <= a) lo = mid + 1;
        else hi = mid;
    }
    counts[r] += M - lo;
}

// Ranking without sorting the nulls: with the R region areas sorted in DEcreasing order (D), a null v
// exceeds exactly the regions at positions >= i(v) = #{D >= v}; a histogram of i(v) over all nulls
// and an inclusive prefix sum give #{null > D[i]} for every region.  O(M log R) searches in an array
// that lives in L1/L2 instead of a radix sort of M keys.
__global__ void __launch_bounds__(256) null_hist_kernel(const double* __restrict__ nulls, int64_t M,
                                                        const double* __restrict__ desc, int64_t R,
                                                        uint32_t* __restrict__ hist) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += (int64_t)gridDim.x * blockDim.x) {
        const double v = nulls[i];
        int64_t lo = 0, hi = R;  // smallest index with desc[idx] < v (R when none; NaN compares false)
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (desc[mid] < v) hi = mid;
            else lo = mid + 1;
        }
        atomicAdd(&hist[lo], 1u);
    }
}

// counts/p-values in sorted (decreasing-area) order from the exclusive prefix of the histogram
__global__ void hist_pval_kernel(const int64_t* __restrict__ excl, const uint32_t* __restrict__ hist, int64_t R,
                                 double weight, double total, double* __restrict__ p, int64_t* __restrict__ counts,
                                 const int32_t* __restrict__ order) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R) return;
    const int64_t c = excl[i] + (int64_t)hist[i];
    if (p) p[i] = __ddiv_rn(weight * (double)c + 1.0, weight * total + 1.0);
    if (counts) counts[order ? order[i] : i] = c;
}

static int null_hist(dfb_ctx* ctx, const double* nulls_dev, int64_t M, const double* desc_dev, int64_t R,
                     uint32_t* hist_dev) {
    if (M <= 0) return DFB_OK;
    KTimer t(ctx, DFB_K_PVAL);
    null_hist_kernel<<<grid_for(M, 256, ctx->sm_count, 16), 256, 0, ctx->stream>>>(nulls_dev, M, desc_dev, R, hist_dev);
    DFB_CUDA(cudaGetLastError());
    return DFB_OK;
}

// p-values of regions given in decreasing-area order against unsorted nulls
static int pval_by_hist(dfb_ctx* ctx, const double* desc_dev, int64_t R, const double* nulls_dev, int64_t M, int weight,
                        double* p_dev) {
    DevBuf<uint32_t> hist;
    DevBuf<int64_t> excl, tot;
    DFB_TRY(hist.alloc(ctx, (size_t)R + 1));
    DFB_TRY(hist.zero());
    DFB_TRY(excl.alloc(ctx, (size_t)R + 1));
    DFB_TRY(tot.alloc(ctx, 1));
    DFB_TRY(null_hist(ctx, nulls_dev, M, desc_dev, R, hist.p));
    DFB_TRY(exclusive_scan_u32(ctx, hist.p, excl.p, R + 1, tot.p, DFB_K_PVAL));
    {
        KTimer t(ctx, DFB_K_PVAL);
        hist_pval_kernel<<<(unsigned)ceil_div(R, 256), 256, 0, ctx->stream>>>(excl.p, hist.p, R, (double)weight, (double)M,
                                                                             p_dev, nullptr, nullptr);
    }
    DFB_CUDA(cudaGetLastError());
    return DFB_OK;
}

int calc_pval_device(dfb_ctx* ctx, const double* areas_dev, int64_t R, const double* nulls_dev, int64_t M, int weight,
                     double* p_dev, DevBuf<double>* keep_sorted = nullptr) {
    if (R == 0) return DFB_OK;
    DevBuf<double> sorted_local;
    DevBuf<double>& sorted = keep_sorted ? *keep_sorted : sorted_local;
    DFB_TRY(sorted.alloc(ctx, (size_t)std::max<int64_t>(M, 1)));
    if (M > 0) DFB_TRY(sort_doubles(ctx, nulls_dev, sorted.p, M));
    {
        KTimer t(ctx, DFB_K_PVAL);
        pval_kernel<<<(unsigned)ceil_div(R, 256), 256, 0, ctx->stream>>>(areas_dev, R, sorted.p, M, (double)weight,
                                                                        p_dev);
    }
    DFB_CUDA(cudaGetLastError());
    return DFB_OK;
}

// ------------------------------------------------------------------------------------------
// FWER: per-permutation maximum (areas are >= 0, so the IEEE bit pattern orders like the value)

// fullNullSummary$area <- abs(stat) * width  (R/mergeResults.R:231): the genome-wide pool ranks this
// product, which differs from the summed area of calculatePvalues in the last ulps
__global__ void merge_area_kernel(const double* __restrict__ stat, const int32_t* __restrict__ width, int64_t M,
                                  double* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = __dmul_rn(fabs(stat[i]), (double)width[i]);
}

__global__ void perm_max_kernel(const double* __restrict__ area, const int32_t* __restrict__ perm, int64_t M,
                                int64_t B, unsigned long long* __restrict__ bits) {
    // permutation ids arrive sorted, so a warp usually sees ONE id: reduce in the warp, one atomic
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t b = -1;
    unsigned long long key = 0ull;  // 0 = no region
    if (i < M) {
        b = (int64_t)perm[i] - 1;
        if (b >= 0 && b < B) key = (unsigned long long)__double_as_longlong(fabs(area[i])) + 1ull;
        else b = -1;
    }
    const int64_t b0 = __shfl_sync(0xffffffffu, b, 0);
    if (__all_sync(0xffffffffu, b == b0 || b < 0)) {
        unsigned long long m = key;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long t = __shfl_xor_sync(0xffffffffu, m, o);
            m = t > m ? t : m;
        }
        // lane 0 may be out of range (b0 < 0) while others are valid: take any valid id
        const unsigned valid = __ballot_sync(0xffffffffu, b >= 0);
        if (valid) {
            const int64_t bb = __shfl_sync(0xffffffffu, b, __ffs(valid) - 1);
            const bool uniform = __all_sync(0xffffffffu, b == bb || b < 0);
            if (uniform) {
                if ((threadIdx.x & 31) == 0 && m) atomicMax(&bits[bb], m);
                return;
            }
        } else {
            return;
        }
    }
    if (b >= 0 && key) atomicMax(&bits[b], key);
}

__global__ void perm_max_decode_kernel(const unsigned long long* __restrict__ bits, int64_t B, double fill,
                                       double* __restrict__ mx) {
    int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const unsigned long long v = bits[b];
    mx[b] = v == 0ull ? fill : __longlong_as_double((long long)(v - 1ull));
}

__global__ void fwer_kernel(const double* __restrict__ areas, int64_t R, const double* __restrict__ mx, int64_t B,
                            double* __restrict__ fwer) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const double a = areas[r];
    int64_t c = 0;
    for (int64_t b = 0; b < B; ++b) c += mx[b] > a ? 1 : 0;
    fwer[r] = __ddiv_rn((double)(c + 1), (double)(B + 1));
}

// ---- multi-GPU pooling (R/mergeResults.R:216-235, 300-307, 380-386).  Per-rank work does not grow
// with the number of GPUs: ranks exchange their REGION areas (small) and a histogram, never the nulls.

// record of a rank: [R | M | max null area of permutation 1..B (-1: none) | region areas, descending, -1 padded
// to rcap].  A rank with more than rcap regions writes the first rcap: every rank sees R > rcap in the
// gathered records and the caller repeats the exchange with a larger capacity.
__global__ void pool_record_kernel(const double* __restrict__ desc, int64_t R, int64_t M, int64_t rcap, int64_t B,
                                   double* __restrict__ rec) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < rcap; i += (int64_t)gridDim.x * blockDim.x)
        rec[2 + B + i] = i < R ? desc[i] : -1.0;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        rec[0] = (double)R;
        rec[1] = (double)M;
    }
}

// Stable merge of the ranks' descending lists without a sort: element i of list w goes to position
// i + sum_{w' < w} #{list w' >= a} + sum_{w' > w} #{list w' > a}  (one binary search per other list).
__global__ void __launch_bounds__(256) pool_merge_kernel(const double* __restrict__ recs, int64_t stride, int64_t off,
                                                         int64_t rcap, int world, double* __restrict__ merged,
                                                         int32_t* __restrict__ src) {
    const int64_t n = (int64_t)world * rcap;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t w = t / rcap, i = t - w * rcap;
        const double a = recs[w * stride + off + i];
        int64_t pos = i;
        for (int w2 = 0; w2 < world; ++w2) {
            if (w2 == w) continue;
            const double* __restrict__ list = recs + (size_t)w2 * stride + off;
            int64_t lo = 0, hi = rcap;
            if (w2 < w) {  // first index with list[idx] < a
                while (lo < hi) {
                    const int64_t mid = (lo + hi) >> 1;
                    if (list[mid] < a) hi = mid;
                    else lo = mid + 1;
                }
            } else {  // first index with list[idx] <= a
                while (lo < hi) {
                    const int64_t mid = (lo + hi) >> 1;
                    if (list[mid] <= a) hi = mid;
                    else lo = mid + 1;
                }
            }
            pos += lo;
        }
        merged[pos] = a;
        src[pos] = (int32_t)t;
    }
}

// i(v) = #{D >= v} over the merged, descending region list D of every rank; 64-bit bins (they are
// summed over ranks and a genome-wide run can exceed 2^32 nulls)
__global__ void __launch_bounds__(256) null_hist64_kernel(const double* __restrict__ nulls, int64_t M,
                                                          const double* __restrict__ desc, int64_t R,
                                                          unsigned long long* __restrict__ hist) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += (int64_t)gridDim.x * blockDim.x) {
        const double v = nulls[i];
        int64_t lo = 0, hi = R;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (desc[mid] < v) hi = mid;
            else lo = mid + 1;
        }
        atomicAdd(&hist[lo], 1ull);
    }
}

// genome-wide maximum area per permutation over all ranks' records; a permutation without a null
// region anywhere contributes cutoffFstatUsed (:382)
__global__ void pool_gmax_kernel(const double* __restrict__ recs, int64_t stride, int world, int64_t B,
                                 double cutoff_used, double* __restrict__ gmax) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double m = -1.0;
    for (int w = 0; w < world; ++w) m = fmax(m, recs[(size_t)w * stride + 2 + b]);
    gmax[b] = m < 0.0 ? cutoff_used : m;
}

// merged position i belongs to region src[i] = rank * rcap + (position in that rank's sorted list);
// this rank's entries get their pooled p-value and FWER, written in the caller's (up-then-dn) order
__global__ void pool_rank_kernel(const unsigned long long* __restrict__ incl, const double* __restrict__ merged,
                                 const int32_t* __restrict__ src, int64_t n, int64_t rcap, int rank, int64_t R,
                                 const int32_t* __restrict__ order, const double* __restrict__ recs, int64_t stride,
                                 int world, const double* __restrict__ gmax, int64_t B, double* __restrict__ p,
                                 double* __restrict__ fwer) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t s = src[i];
    const int64_t w = s / rcap, il = s - w * rcap;
    if (w != rank || il >= R) return;
    double total = 0.0;  // pooled null regions (each is stored twice by the reference: weight 2 below)
    for (int w2 = 0; w2 < world; ++w2) total += recs[(size_t)w2 * stride + 1];
    const int64_t dst = order ? order[il] : il;
    p[dst] = __ddiv_rn(2.0 * (double)incl[i] + 1.0, 2.0 * total + 1.0);
    const double a = merged[i];
    int64_t k = 0;
    for (int64_t b = 0; b < B; ++b) k += gmax[b] > a ? 1 : 0;
    fwer[dst] = __ddiv_rn((double)(k + 1), (double)(B + 1));
}

int perm_max_device(dfb_ctx* ctx, const double* area_dev, const int32_t* perm_dev, int64_t M, int64_t B, double fill,
                    double* mx_dev) {
    DevBuf<unsigned long long> bits;
    DFB_TRY(bits.alloc(ctx, (size_t)std::max<int64_t>(B, 1)));
    DFB_TRY(bits.zero());
    KTimer t(ctx, DFB_K_PVAL, 2);
    if (M > 0)
        perm_max_kernel<<<(unsigned)ceil_div(M, 256), 256, 0, ctx->stream>>>(area_dev, perm_dev, M, B, bits.p);
    if (B > 0) perm_max_decode_kernel<<<(unsigned)ceil_div(B, 256), 256, 0, ctx->stream>>>(bits.p, B, fill, mx_dev);
    DFB_CUDA(cudaGetLastError());
    return DFB_OK;
}

// ------------------------------------------------------------------------------------------
// quantile type 7 (na.rm = TRUE)

// Order statistics WITHOUT sorting: radix select over the order-preserving 64-bit image of the doubles, 11
// bits per pass (2048 shared-memory bins), six read-only passes over x, all bookkeeping on the device (one
// host synchronisation at the end).  The default cutoff quantile(fstats, 0.99) (R/findRegions.R:118,
// R/calculatePvalues.R:160, R/analyzeChr.R:318) used to be a full radix sort of the N statistics.
constexpr int QS_BITS = 11, QS_BINS = 1 << QS_BITS, QS_PASSES = 6;  // 6 * 11 = 66 >= 64

__device__ __forceinline__ unsigned long long qs_key(double v) {  // ascending order of doubles == ascending order of keys
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double qs_unkey(unsigned long long k) {
    const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}

// state[0] = key prefix fixed so far, state[1] = rank still to find inside the prefix (0-based), state[2] = NaN count,
// state[3] = number of non-NaN values m
__global__ void __launch_bounds__(256) qs_hist_kernel(const double* __restrict__ x, int64_t N, int pass,
                                                      const unsigned long long* __restrict__ state,
                                                      unsigned int* __restrict__ hist, unsigned long long* __restrict__ nan_count) {
    __shared__ unsigned int h[QS_BINS];
    for (int i = threadIdx.x; i < QS_BINS; i += blockDim.x) h[i] = 0u;
    __syncthreads();
    const int shift = 64 - QS_BITS * (pass + 1);  // negative in the last pass: the remaining low bits
    const unsigned long long prefix = pass ? state[0] : 0ull;
    unsigned long long nn = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        const double v = x[i];
        if (v != v) {
            ++nn;
            continue;
        }
        const unsigned long long k = qs_key(v);
        const int hs = 64 - QS_BITS * pass;  // bits above the current digit
        if (pass && (hs >= 64 ? 0ull : (k >> hs)) != prefix) continue;
        const unsigned int d = shift >= 0 ? (unsigned int)((k >> shift) & (QS_BINS - 1)) : (unsigned int)((k << (-shift)) & (QS_BINS - 1));
        atomicAdd(&h[d], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < QS_BINS; i += blockDim.x)
        if (h[i]) atomicAdd(&hist[i], h[i]);
    if (pass == 0 && nan_count) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nn += __shfl_xor_sync(0xffffffffu, nn, o);
        if ((threadIdx.x & 31) == 0 && nn) atomicAdd(nan_count, nn);
    }
}

// one block: pick the bin that holds the wanted rank, extend the prefix, clear the histogram for the next pass
__global__ void __launch_bounds__(QS_BINS / 2) qs_pick_kernel(unsigned int* __restrict__ hist, int pass, int64_t N, double prob,
                                                              int which, unsigned long long* __restrict__ state) {
    __shared__ unsigned long long cum[QS_BINS];
    __shared__ unsigned long long want;
    if (threadIdx.x == 0) {
        if (pass == 0) {
            const unsigned long long m = (unsigned long long)N - state[2];
            state[3] = m;
            // stats::quantile type 7: index = 1 + (m - 1) * p; ranks floor(index) and ceil(index) (1-based)
            const double index = 1.0 + (double)(m ? m - 1 : 0) * prob;
            const double r = which == 0 ? floor(index) : ceil(index);
            state[1] = (unsigned long long)(r - 1.0);
            state[0] = 0ull;
        }
        want = state[1];
    }
    __shared__ unsigned long long part[64];
    for (int i = threadIdx.x; i < QS_BINS; i += blockDim.x) cum[i] = hist[i];
    __syncthreads();
    if (threadIdx.x < 64) {  // 64 partial sums of 32 bins, then two short serial scans
        unsigned long long sum = 0;
        for (int i = 0; i < 32; ++i) sum += cum[threadIdx.x * 32 + i];
        part[threadIdx.x] = sum;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long run = 0;
        int blk = 63;
        for (int i = 0; i < 64; ++i) {
            if (want < run + part[i]) {
                blk = i;
                break;
            }
            run += part[i];
        }
        int pick = blk * 32 + 31;
        for (int i = blk * 32; i < blk * 32 + 32; ++i) {
            if (want < run + cum[i]) {
                pick = i;
                break;
            }
            run += cum[i];
        }
        state[1] = want - run;
        const int bits = (64 - QS_BITS * pass) >= QS_BITS ? QS_BITS : (64 - QS_BITS * pass);  // 9 bits in the last pass
        state[0] = (state[0] << bits) | (unsigned long long)(bits == QS_BITS ? pick : (pick >> (QS_BITS - bits)));
    }
    __syncthreads();
    for (int i = threadIdx.x; i < QS_BINS; i += blockDim.x) hist[i] = 0u;
}

__global__ void qs_result_kernel(const unsigned long long* __restrict__ state, double* __restrict__ out) {
    out[0] = qs_unkey(state[0]);
}

// k-th smallest (type-7 rank `which`: 0 = floor(index), 1 = ceil(index)) of the non-NaN values of x; *m_out = their count
static int select_rank(dfb_ctx* ctx, const double* x_dev, int64_t N, double prob, int which, double* value_dev,
                       unsigned long long* state_dev, unsigned int* hist_dev) {
    const int grid = grid_for(N, 256, ctx->sm_count, 8);
    KTimer t(ctx, DFB_K_PVAL, 2 * QS_PASSES + 1);
    for (int pass = 0; pass < QS_PASSES; ++pass) {
        qs_hist_kernel<<<grid, 256, 0, ctx->stream>>>(x_dev, N, pass, state_dev, hist_dev, pass == 0 && which == 0 ? state_dev + 2 : nullptr);
        qs_pick_kernel<<<1, QS_BINS / 2, 0, ctx->stream>>>(hist_dev, pass, N, prob, which, state_dev);
    }
    qs_result_kernel<<<1, 1, 0, ctx->stream>>>(state_dev, value_dev);
    DFB_CUDA(cudaGetLastError());
    return DFB_OK;
}

int quantile_device(dfb_ctx* ctx, const double* x_dev, int64_t N, double prob, double* out) {
    DFB_REQUIRE(prob >= 0 && prob <= 1, "'probs' outside [0,1]");
    if (N <= 0) {
        *out = nan("");
        return DFB_OK;
    }
    DevBuf<unsigned long long> state;
    DevBuf<unsigned int> hist;
    DevBuf<double> vals;
    DFB_TRY(state.alloc(ctx, 4));
    DFB_TRY(hist.alloc(ctx, QS_BINS));
    DFB_TRY(vals.alloc(ctx, 2));
    DFB_TRY(state.zero());
    DFB_TRY(hist.zero());
    DFB_TRY(select_rank(ctx, x_dev, N, prob, 0, vals.p, state.p, hist.p));      // x[floor(index)]
    DFB_TRY(select_rank(ctx, x_dev, N, prob, 1, vals.p + 1, state.p, hist.p));  // x[ceil(index)] (NaN count kept in state[2])
    double v[2];
    unsigned long long st[4];
    DFB_TRY(d2h(ctx, v, vals.p, sizeof v));
    DFB_TRY(d2h(ctx, st, state.p, sizeof st));
    const int64_t m = (int64_t)st[3];
    if (m <= 0) {
        *out = nan("");
        return DFB_OK;
    }
    // stats::quantile type 7: qs = x[lo]; if index > lo and x[hi] != qs: qs = (1-h)*qs + h*x[hi]
    const double index = 1.0 + (double)(m - 1) * prob;
    const double lo = floor(index);
    double qs = v[0];
    if (index > lo && v[1] != qs) {
        const double h = index - lo;
        qs = (1.0 - h) * qs + h * v[1];
    }
    *out = qs;
    return DFB_OK;
}

__global__ void iota_kernel(int32_t* __restrict__ x, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = (int32_t)i;
}

// stable order(area, decreasing = TRUE): descending radix sort of (area, iota) pairs is stable
static int order_desc(dfb_ctx* ctx, const double* area_dev, int64_t R, int32_t* order_dev, double* sorted_dev) {
    DevBuf<int32_t> iota;
    DFB_TRY(iota.alloc(ctx, (size_t)R));
    {
        KTimer t(ctx, DFB_K_PVAL);
        iota_kernel<<<(unsigned)ceil_div(R, 256), 256, 0, ctx->stream>>>(iota.p, R);
    }
    size_t tmp_bytes = 0;
    DFB_CUDA(cub::DeviceRadixSort::SortPairsDescending(nullptr, tmp_bytes, area_dev, sorted_dev, iota.p, order_dev, R, 0,
                                                       64, ctx->stream));
    DevBuf<unsigned char> tmp;
    DFB_TRY(tmp.alloc(ctx, tmp_bytes));
    {
        KTimer t(ctx, DFB_K_PVAL, 8);
        DFB_CUDA(cub::DeviceRadixSort::SortPairsDescending(tmp.p, tmp_bytes, area_dev, sorted_dev, iota.p, order_dev, R,
                                                           0, 64, ctx->stream));
    }
    return DFB_OK;  // stream-ordered: temporaries are released after the sort
}

}  // namespace dfb

using namespace dfb;

extern "C" {

int dfb_fstats(dfb_ctx* ctx, dfb_cov* cov, const double* mod, int p, const double* mod0, int p0, double adjust_f,
               double* fstats_out) {
    DFB_REQUIRE(ctx && cov, "dfb_fstats: NULL argument");
    if (!cov->y.p && !cov->y_lazy) {
        set_error("preprocessCoverage() has not been run on this coverage (coverageProcessed is missing)");
        return DFB_ESTATE;
    }
    ModelBasis mb;
    DFB_TRY(build_model_basis(mod, cov->n, p, mod0, p0, &mb));
    DFB_TRY(cov->fstats.alloc(ctx, (size_t)std::max<int64_t>(cov->N, 1)));
    cov->has_f = false;
    if (cov->N > 0) DFB_TRY(fstat_observed(ctx, cov, mb, adjust_f, cov->fstats.p));
    cov->has_f = true;
    // without a host destination the statistics stay on the device and the call returns without
    // synchronising: everything it enqueued is stream-ordered (the basis was staged by the copy)
    if (fstats_out) DFB_TRY(d2h_big(ctx, fstats_out, cov->fstats.p, (size_t)cov->N * sizeof(double)));
    return DFB_OK;
}

int dfb_fstats_begin(dfb_ctx* ctx, dfb_cov* cov, const double* mod, int p, const double* mod0, int p0, double adjust_f,
                     double* fstats_out) {
    DFB_REQUIRE(fstats_out, "dfb_fstats_begin: NULL destination (use dfb_fstats to keep the statistics on the device)");
    DFB_TRY(dfb_fstats(ctx, cov, mod, p, mod0, p0, adjust_f, nullptr));
    return bg_download_begin(ctx, fstats_out, cov->fstats.p, (size_t)cov->N * sizeof(double));
}

int dfb_fstats_wait(dfb_ctx* ctx) {
    DFB_REQUIRE(ctx, "dfb_fstats_wait: NULL context");
    return bg_download_wait(ctx);
}

int dfb_perm_nulls(dfb_ctx* ctx, dfb_cov* cov, const double* mod, int p, const double* mod0, int p0, double adjust_f,
                   const int32_t* perm_idx, int64_t B, double cutoff_lo, double cutoff_hi, int32_t max_region_gap,
                   dfb_nulls** out) {
    DFB_REQUIRE(ctx && cov && out, "dfb_perm_nulls: NULL argument");
    arena_reset(ctx);
    return perm_nulls_impl(ctx, cov, mod, p, mod0, p0, adjust_f, perm_idx, B, cutoff_lo, cutoff_hi, max_region_gap,
                           out);
}

void dfb_nulls_free(dfb_nulls* nl) {
    if (!nl) return;
    cudaStreamSynchronize(nl->ctx->stream);
    delete nl;
}

int64_t dfb_nulls_count(const dfb_nulls* nl) { return nl ? nl->count : 0; }

int dfb_nulls_counts_per_perm(const dfb_nulls* nl, int64_t* counts) {
    DFB_REQUIRE(nl && counts, "dfb_nulls_counts_per_perm: NULL argument");
    memcpy(counts, nl->per_perm.data(), nl->per_perm.size() * sizeof(int64_t));
    return DFB_OK;
}

int dfb_nulls_get(const dfb_nulls* nl, int dup, double* stat, int32_t* width, double* area, int32_t* permutation) {
    DFB_REQUIRE(nl, "dfb_nulls_get: NULL handle");
    const size_t M = (size_t)nl->count;
    if (M == 0) return DFB_OK;
    dfb_ctx* ctx = nl->ctx;
    if (!dup) {
        if (stat) DFB_TRY(d2h_big(ctx, stat, nl->stat.p, M * 8));
        if (area) DFB_TRY(d2h_big(ctx, area, nl->area.p, M * 8));
        if (width) DFB_TRY(d2h_big(ctx, width, nl->width.p, M * 4));
        if (permutation) DFB_TRY(d2h_big(ctx, permutation, nl->perm.p, M * 4));
        return DFB_OK;
    }
    // the reference appends every permutation's regions twice: [r1..rk, r1..rk] (:346-354).  Every
    // array crosses PCIe once; the duplicated layout is written by the host copy pool while the
    // next chunk is in flight.
    const int64_t B = (int64_t)nl->per_perm.size();
    std::vector<int64_t> pre((size_t)B + 1, 0);
    for (int64_t b = 0; b < B; ++b) pre[(size_t)b + 1] = pre[(size_t)b] + nl->per_perm[(size_t)b];
    DFB_REQUIRE(pre[(size_t)B] == (int64_t)M, "internal error: per-permutation counts do not add up");
    void* dsts[4];
    const void* srcs[4];
    size_t elems[4];
    int na = 0;
    auto add = [&](void* d, const void* sp, size_t e) {
        if (d) {
            dsts[na] = d;
            srcs[na] = sp;
            elems[na] = e;
            ++na;
        }
    };
    add(stat, nl->stat.p, 8);
    add(area, nl->area.p, 8);
    add(width, nl->width.p, 4);
    add(permutation, nl->perm.p, 4);
    return d2h_dup_multi(ctx, na, dsts, srcs, elems, pre.data(), B);
}

int dfb_nulls_copy_dev(const dfb_nulls* nl, double* area_dev, int32_t* perm_dev) {
    DFB_REQUIRE(nl, "dfb_nulls_copy_dev: NULL handle");
    const size_t M = (size_t)nl->count;
    if (M == 0) return DFB_OK;
    if (area_dev) DFB_CUDA(cudaMemcpyAsync(area_dev, nl->area.p, M * 8, cudaMemcpyDeviceToDevice, nl->ctx->stream));
    if (perm_dev) DFB_CUDA(cudaMemcpyAsync(perm_dev, nl->perm.p, M * 4, cudaMemcpyDeviceToDevice, nl->ctx->stream));
    return DFB_OK;
}

int dfb_nulls_merge_areas_dev(const dfb_nulls* nl, double* area_dev) {
    DFB_REQUIRE(nl && (area_dev || nl->count == 0), "dfb_nulls_merge_areas_dev: NULL argument");
    if (nl->count == 0) return DFB_OK;
    dfb_ctx* ctx = nl->ctx;
    {
        KTimer t(ctx, DFB_K_PVAL);
        merge_area_kernel<<<(unsigned)std::min<int64_t>(ceil_div(nl->count, 256), (int64_t)ctx->sm_count * 8), 256, 0,
                            ctx->stream>>>(nl->stat.p, nl->width.p, nl->count, area_dev);
    }
    DFB_CUDA(cudaGetLastError());
    return DFB_OK;
}

static int pooled_null_areas(dfb_ctx* ctx, dfb_nulls* nl, int merge_area, const double** out) {
    *out = nl->area.p;
    if (!merge_area || nl->count == 0) return DFB_OK;
    if (!nl->area_merge.p) {
        DFB_TRY(nl->area_merge.alloc(ctx, (size_t)nl->count));
        KTimer t(ctx, DFB_K_PVAL);
        merge_area_kernel<<<(unsigned)std::min<int64_t>(ceil_div(nl->count, 256), (int64_t)ctx->sm_count * 8), 256, 0,
                            ctx->stream>>>(nl->stat.p, nl->width.p, nl->count, nl->area_merge.p);
        DFB_CUDA(cudaGetLastError());
    }
    *out = nl->area_merge.p;
    return DFB_OK;
}

int dfb_pool_local_dev(dfb_ctx* ctx, const dfb_regions* r, dfb_nulls* nl, int64_t B, int merge_area, int64_t rcap,
                       double* record_dev) {
    DFB_REQUIRE(ctx && record_dev, "dfb_pool_local_dev: NULL argument");
    DFB_REQUIRE(B >= 1 && rcap >= 1, "dfb_pool_local_dev: nPermute and capacity must be >= 1");
    if (!r || !nl) {
        // a chromosome without observed regions: calculatePvalues() returned list(regions = NULL, nullStats =
        // NULL, ...) before its permutation loop (:245-251), so it adds nothing to the genome-wide pool
        DFB_REQUIRE(!r && !nl, "dfb_pool_local_dev: regions and nulls must both be NULL for an empty chromosome");
        DevBuf<double> none;
        DFB_TRY(none.alloc(ctx, 1));
        {
            KTimer t(ctx, DFB_K_PVAL);
            pool_record_kernel<<<grid_for(rcap, 256, ctx->sm_count, 8), 256, 0, ctx->stream>>>(none.p, 0, 0, rcap, B,
                                                                                          record_dev);
        }
        DFB_CUDA(cudaGetLastError());
        return perm_max_device(ctx, nullptr, nullptr, 0, B, -1.0, record_dev + 2);
    }
    if (r->R > 0 && !r->d_area_desc.p) {
        set_error("dfb_pool_local_dev: this result was not produced by dfb_calculate_pvalues");
        return DFB_ESTATE;
    }
    const double* areas = nullptr;
    DFB_TRY(pooled_null_areas(ctx, nl, merge_area, &areas));
    {
        KTimer t(ctx, DFB_K_PVAL);
        pool_record_kernel<<<grid_for(rcap, 256, ctx->sm_count, 8), 256, 0, ctx->stream>>>(r->d_area_desc.p, r->R,
                                                                                      nl->count, rcap, B, record_dev);
    }
    DFB_CUDA(cudaGetLastError());
    return perm_max_device(ctx, areas, nl->perm.p, nl->count, B, -1.0, record_dev + 2);
}

int dfb_pool_hist_dev(dfb_ctx* ctx, dfb_nulls* nl, int merge_area, const double* records_dev, int world, int64_t B,
                      int64_t rcap, double* merged_dev, int32_t* src_dev, unsigned long long* hist_dev) {
    DFB_REQUIRE(ctx && records_dev && merged_dev && src_dev && hist_dev, "dfb_pool_hist_dev: NULL argument");
    DFB_REQUIRE(world >= 1 && B >= 1 && rcap >= 1, "dfb_pool_hist_dev: world, nPermute and capacity must be >= 1");
    const int64_t n = (int64_t)world * rcap;
    DFB_REQUIRE(n < INT32_MAX, "dfb_pool_hist_dev: %lld pooled regions exceed the 32-bit index", (long long)n);
    {
        KTimer t(ctx, DFB_K_PVAL);
        pool_merge_kernel<<<grid_for(n, 256, ctx->sm_count, 8), 256, 0, ctx->stream>>>(records_dev, 2 + B + rcap, 2 + B,
                                                                                    rcap, world, merged_dev, src_dev);
    }
    DFB_CUDA(cudaGetLastError());
    DFB_CUDA(cudaMemsetAsync(hist_dev, 0, (size_t)(n + 1) * sizeof(unsigned long long), ctx->stream));
    const double* areas = nullptr;
    if (nl) DFB_TRY(pooled_null_areas(ctx, nl, merge_area, &areas));
    if (nl && nl->count > 0) {
        KTimer t(ctx, DFB_K_PVAL);
        null_hist64_kernel<<<grid_for(nl->count, 256, ctx->sm_count, 16), 256, 0, ctx->stream>>>(areas, nl->count,
                                                                                             merged_dev, n, hist_dev);
    }
    DFB_CUDA(cudaGetLastError());
    return DFB_OK;
}

int dfb_pool_rank_dev(dfb_ctx* ctx, const dfb_regions* r, const double* records_dev, int world, int rank, int64_t B,
                      int64_t rcap, const double* merged_dev, const int32_t* src_dev,
                      const unsigned long long* hist_dev, double cutoff_used, double* p_dev, double* fwer_dev) {
    DFB_REQUIRE(ctx && records_dev && merged_dev && src_dev && hist_dev, "dfb_pool_rank_dev: NULL argument");
    DFB_REQUIRE(world >= 1 && rank >= 0 && rank < world && B >= 1 && rcap >= 1,
                "dfb_pool_rank_dev: bad world / rank / capacity");
    if (!r) return DFB_OK;  // empty chromosome: nothing to rank
    const int64_t R = std::min<int64_t>(r->R, rcap);  // R > rcap: the caller repeats with a larger capacity
    if (R == 0) return DFB_OK;
    DFB_REQUIRE(p_dev && fwer_dev, "dfb_pool_rank_dev: NULL output");
    const int64_t n = (int64_t)world * rcap;
    DevBuf<unsigned long long> incl;
    DevBuf<double> gmax;
    DevBuf<unsigned char> tmp;
    DFB_TRY(incl.alloc(ctx, (size_t)n));
    DFB_TRY(gmax.alloc(ctx, (size_t)B));
    size_t tmp_bytes = 0;
    DFB_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tmp_bytes, hist_dev, incl.p, n, ctx->stream));
    DFB_TRY(tmp.alloc(ctx, tmp_bytes));
    {
        KTimer t(ctx, DFB_K_PVAL, 4);
        DFB_CUDA(cub::DeviceScan::InclusiveSum(tmp.p, tmp_bytes, hist_dev, incl.p, n, ctx->stream));
        pool_gmax_kernel<<<(unsigned)ceil_div(B, 256), 256, 0, ctx->stream>>>(records_dev, 2 + B + rcap, world, B,
                                                                             cutoff_used, gmax.p);
        pool_rank_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, ctx->stream>>>(incl.p, merged_dev, src_dev, n, rcap, rank, R,
                                                                            r->d_order.p, records_dev, 2 + B + rcap,
                                                                            world, gmax.p, B, p_dev, fwer_dev);
    }
    DFB_CUDA(cudaGetLastError());
    return DFB_OK;
}

// scatter of the pooled results from the rank's merged (decreasing-area) order back to chromosome order
__global__ void merge_scatter_kernel(const double* __restrict__ p_sorted, const double* __restrict__ f_sorted,
                                     const int32_t* __restrict__ payload, int64_t R, double* __restrict__ p_cat,
                                     double* __restrict__ f_cat) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R) return;
    p_cat[payload[i]] = p_sorted[i];
    f_cat[payload[i]] = f_sorted[i];
}

__global__ void merge_header_kernel(double* __restrict__ rec, double R, double M) {
    rec[0] = R;
    rec[1] = M;
}

int dfb_merge_results(dfb_ctx* ctx, int nchr, dfb_regions* const* r, dfb_nulls* const* nl, int64_t B,
                      double cutoff_used, int merge_area, double* const* pvalues, double* const* fwer,
                      int64_t* total_nulls) {
    DFB_REQUIRE(ctx && nchr >= 0 && (nchr == 0 || (r && nl)) && B >= 1, "dfb_merge_results: bad argument");
    arena_reset(ctx);
    const int world = ctx->world, rank = ctx->rank;
    // ---- this rank's chromosomes: regions concatenated (each list is already in decreasing-area order)
    std::vector<int64_t> roff((size_t)nchr + 1, 0);
    int64_t M_loc = 0;
    for (int c = 0; c < nchr; ++c) {
        DFB_REQUIRE((r[c] == nullptr) == (nl[c] == nullptr), "dfb_merge_results: chromosome %d has only one of regions / nulls", c);
        const int64_t Rc = r[c] ? r[c]->R : 0;
        if (Rc > 0 && !r[c]->d_area_desc.p) {
            set_error("dfb_merge_results: chromosome %d was not produced by dfb_calculate_pvalues", c);
            return DFB_ESTATE;
        }
        roff[(size_t)c + 1] = roff[(size_t)c] + Rc;
        if (nl[c]) {
            DFB_REQUIRE(nl[c]->B == B, "dfb_merge_results: chromosome %d was analysed with %lld permutations, not %lld", c,
                        (long long)nl[c]->B, (long long)B);
            M_loc += nl[c]->count;
        }
    }
    const int64_t R_loc = roff[(size_t)nchr];
    DFB_REQUIRE(R_loc < INT32_MAX, "dfb_merge_results: too many regions on one rank");
    // ---- every rank learns every rank's counts (the one host synchronisation of the exchange)
    DevBuf<int64_t> cnt_d;
    DFB_TRY(cnt_d.alloc(ctx, (size_t)2 * (world + 1)));
    const int64_t mine[2] = {R_loc, M_loc};
    DFB_TRY(h2d(ctx, cnt_d.p + 2 * world, mine, sizeof mine));
    DFB_TRY(comm_all_gather(ctx, cnt_d.p + 2 * world, cnt_d.p, 2, sizeof(int64_t)));
    std::vector<int64_t> cnt((size_t)2 * world);
    DFB_TRY(d2h(ctx, cnt.data(), cnt_d.p, cnt.size() * sizeof(int64_t)));
    int64_t rcap = 1, M_tot = 0;
    for (int w = 0; w < world; ++w) {
        rcap = std::max(rcap, cnt[(size_t)2 * w]);
        M_tot += cnt[(size_t)2 * w + 1];
    }
    if (total_nulls) *total_nulls = M_tot;
    const int64_t n = (int64_t)world * rcap;
    DFB_REQUIRE(n < INT32_MAX, "dfb_merge_results: %lld pooled regions exceed the 32-bit index", (long long)n);
    // ---- local merged list (sort of the concatenation; payload = position in the concatenation)
    DevBuf<double> cat, desc;
    DevBuf<int32_t> payload;
    DFB_TRY(cat.alloc(ctx, (size_t)std::max<int64_t>(R_loc, 1)));
    DFB_TRY(desc.alloc(ctx, (size_t)std::max<int64_t>(R_loc, 1)));
    DFB_TRY(payload.alloc(ctx, (size_t)std::max<int64_t>(R_loc, 1)));
    for (int c = 0; c < nchr; ++c)
        if (r[c] && r[c]->R > 0)
            DFB_CUDA(cudaMemcpyAsync(cat.p + roff[(size_t)c], r[c]->d_area_desc.p, (size_t)r[c]->R * 8, cudaMemcpyDeviceToDevice,
                                     ctx->stream));
    if (R_loc > 0) DFB_TRY(order_desc(ctx, cat.p, R_loc, payload.p, desc.p));
    // ---- record {R, M, per-permutation max null area (-1: none), region areas desc padded with -1}
    const int64_t stride = 2 + B + rcap;
    DevBuf<double> recs;
    DFB_TRY(recs.alloc(ctx, (size_t)(world + 1) * stride));
    double* rec = recs.p + (size_t)world * stride;
    {
        KTimer t(ctx, DFB_K_PVAL);
        pool_record_kernel<<<grid_for(rcap, 256, ctx->sm_count, 8), 256, 0, ctx->stream>>>(desc.p, R_loc, M_loc, rcap, B, rec);
    }
    {
        DevBuf<unsigned long long> bits;
        DFB_TRY(bits.alloc(ctx, (size_t)B));
        DFB_TRY(bits.zero());
        KTimer t(ctx, DFB_K_PVAL, nchr + 1);
        for (int c = 0; c < nchr; ++c) {
            if (!nl[c] || nl[c]->count == 0) continue;
            const double* areas = nullptr;
            DFB_TRY(pooled_null_areas(ctx, nl[c], merge_area, &areas));
            perm_max_kernel<<<(unsigned)ceil_div(nl[c]->count, 256), 256, 0, ctx->stream>>>(areas, nl[c]->perm.p, nl[c]->count,
                                                                                           B, bits.p);
        }
        perm_max_decode_kernel<<<(unsigned)ceil_div(B, 256), 256, 0, ctx->stream>>>(bits.p, B, -1.0, rec + 2);
        DFB_CUDA(cudaGetLastError());
    }
    DFB_TRY(comm_all_gather(ctx, rec, recs.p, (size_t)stride, sizeof(double)));
    // ---- merged region list of all ranks, histogram of this rank's nulls over it, summed over the ranks
    DevBuf<double> merged;
    DevBuf<int32_t> src;
    DevBuf<unsigned long long> hist, incl;
    DFB_TRY(merged.alloc(ctx, (size_t)n));
    DFB_TRY(src.alloc(ctx, (size_t)n));
    DFB_TRY(hist.alloc(ctx, (size_t)n + 1));
    DFB_TRY(incl.alloc(ctx, (size_t)n));
    DFB_TRY(hist.zero());
    {
        KTimer t(ctx, DFB_K_PVAL);
        pool_merge_kernel<<<grid_for(n, 256, ctx->sm_count, 8), 256, 0, ctx->stream>>>(recs.p, stride, 2 + B, rcap, world,
                                                                                    merged.p, src.p);
    }
    for (int c = 0; c < nchr; ++c) {
        if (!nl[c] || nl[c]->count == 0) continue;
        const double* areas = nullptr;
        DFB_TRY(pooled_null_areas(ctx, nl[c], merge_area, &areas));
        KTimer t(ctx, DFB_K_PVAL);
        null_hist64_kernel<<<grid_for(nl[c]->count, 256, ctx->sm_count, 16), 256, 0, ctx->stream>>>(areas, nl[c]->count,
                                                                                                merged.p, n, hist.p);
    }
    DFB_CUDA(cudaGetLastError());
    DFB_TRY(comm_all_reduce_sum_u64(ctx, hist.p, (size_t)n + 1));
    // ---- pooled p-values and FWER of this rank's regions
    if (R_loc > 0) {
        DevBuf<double> gmax, p_s, f_s, p_cat, f_cat;
        DevBuf<unsigned char> tmp;
        DFB_TRY(gmax.alloc(ctx, (size_t)B));
        DFB_TRY(p_s.alloc(ctx, (size_t)R_loc));
        DFB_TRY(f_s.alloc(ctx, (size_t)R_loc));
        DFB_TRY(p_cat.alloc(ctx, (size_t)R_loc));
        DFB_TRY(f_cat.alloc(ctx, (size_t)R_loc));
        size_t tmp_bytes = 0;
        DFB_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tmp_bytes, hist.p, incl.p, n, ctx->stream));
        DFB_TRY(tmp.alloc(ctx, tmp_bytes));
        {
            KTimer t(ctx, DFB_K_PVAL, 5);
            DFB_CUDA(cub::DeviceScan::InclusiveSum(tmp.p, tmp_bytes, hist.p, incl.p, n, ctx->stream));
            pool_gmax_kernel<<<(unsigned)ceil_div(B, 256), 256, 0, ctx->stream>>>(recs.p, stride, world, B, cutoff_used, gmax.p);
            pool_rank_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, ctx->stream>>>(incl.p, merged.p, src.p, n, rcap, rank, R_loc,
                                                                                nullptr, recs.p, stride, world, gmax.p, B,
                                                                                p_s.p, f_s.p);
            merge_scatter_kernel<<<(unsigned)ceil_div(R_loc, 256), 256, 0, ctx->stream>>>(p_s.p, f_s.p, payload.p, R_loc,
                                                                                         p_cat.p, f_cat.p);
        }
        DFB_CUDA(cudaGetLastError());
        const double* hp = (pvalues) ? d2h_async_pinned(ctx, p_cat.p, (size_t)R_loc) : nullptr;
        const double* hf = (fwer) ? d2h_async_pinned(ctx, f_cat.p, (size_t)R_loc) : nullptr;
        DFB_CUDA(cudaStreamSynchronize(ctx->stream));
        DFB_REQUIRE((!pvalues || hp) && (!fwer || hf), "dfb_merge_results: read-back failed");
        for (int c = 0; c < nchr; ++c) {
            const int64_t Rc = roff[(size_t)c + 1] - roff[(size_t)c];
            if (Rc == 0) continue;
            if (pvalues && pvalues[c]) {
                if (M_tot > 0) memcpy(pvalues[c], hp + roff[(size_t)c], (size_t)Rc * 8);
                else for (int64_t i = 0; i < Rc; ++i) pvalues[c][i] = nan("");  // no nulls anywhere: NA (:408-419)
            }
            if (fwer && fwer[c]) memcpy(fwer[c], hf + roff[(size_t)c], (size_t)Rc * 8);
        }
    } else {
        DFB_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    return DFB_OK;
}

int dfb_nulls_sorted_areas_dev(dfb_ctx* ctx, dfb_nulls* nl, double* out_dev) {
    DFB_REQUIRE(ctx && nl && (out_dev || nl->count == 0), "dfb_nulls_sorted_areas_dev: NULL argument");
    const size_t M = (size_t)nl->count;
    if (M == 0) return DFB_OK;
    if (!nl->area_sorted.p) {  // not produced by dfb_calculate_pvalues (e.g. dfb_perm_nulls): sort now
        DFB_TRY(nl->area_sorted.alloc(ctx, M));
        DFB_TRY(sort_doubles(ctx, nl->area.p, nl->area_sorted.p, (int64_t)M));
    }
    DFB_CUDA(cudaMemcpyAsync(out_dev, nl->area_sorted.p, M * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    return DFB_OK;
}

int dfb_count_greater_dev(dfb_ctx* ctx, const double* areas_dev, int64_t R, const double* sorted_nulls_dev, int64_t M,
                          int64_t* counts_dev) {
    DFB_REQUIRE(ctx && (areas_dev || R == 0) && (sorted_nulls_dev || M == 0) && (counts_dev || R == 0),
                "dfb_count_greater_dev: NULL argument");
    if (R == 0) return DFB_OK;
    {
        KTimer t(ctx, DFB_K_PVAL);
        count_greater_kernel<<<(unsigned)ceil_div(R, 256), 256, 0, ctx->stream>>>(areas_dev, R, sorted_nulls_dev, M,
                                                                                 counts_dev);
    }
    DFB_CUDA(cudaGetLastError());
    return DFB_OK;
}

int dfb_regions_null_hist_dev(dfb_ctx* ctx, const dfb_regions* r, const double* nulls_dev, int64_t M,
                              uint32_t* hist_dev) {
    DFB_REQUIRE(ctx && r && (nulls_dev || M == 0) && hist_dev, "dfb_regions_null_hist_dev: NULL argument");
    if (!r->d_area_desc.p) {
        set_error("dfb_regions_null_hist_dev: this result was not produced by dfb_calculate_pvalues");
        return DFB_ESTATE;
    }
    return null_hist(ctx, nulls_dev, M, r->d_area_desc.p, r->R, hist_dev);
}

int dfb_regions_hist_counts_dev(dfb_ctx* ctx, const dfb_regions* r, const uint32_t* hist_dev, int64_t* counts_dev) {
    DFB_REQUIRE(ctx && r && hist_dev && (counts_dev || r->R == 0), "dfb_regions_hist_counts_dev: NULL argument");
    if (!r->d_area_desc.p) {
        set_error("dfb_regions_hist_counts_dev: this result was not produced by dfb_calculate_pvalues");
        return DFB_ESTATE;
    }
    const int64_t R = r->R;
    if (R == 0) return DFB_OK;
    DevBuf<int64_t> excl, tot;
    DFB_TRY(excl.alloc(ctx, (size_t)R + 1));
    DFB_TRY(tot.alloc(ctx, 1));
    DFB_TRY(exclusive_scan_u32(ctx, hist_dev, excl.p, R + 1, tot.p, DFB_K_PVAL));
    {
        KTimer t(ctx, DFB_K_PVAL);
        hist_pval_kernel<<<(unsigned)ceil_div(R, 256), 256, 0, ctx->stream>>>(excl.p, hist_dev, R, 0.0, 0.0, nullptr,
                                                                             counts_dev, r->d_order.p);
    }
    DFB_CUDA(cudaGetLastError());
    return DFB_OK;
}

int dfb_regions_copy_areas_dev(dfb_ctx* ctx, const dfb_regions* r, double* out_dev) {
    DFB_REQUIRE(ctx && r && (out_dev || r->R == 0), "dfb_regions_copy_areas_dev: NULL argument");
    if (r->R == 0) return DFB_OK;
    if (!r->d_area.p) {
        set_error("dfb_regions_copy_areas_dev: this result holds no device copy of the areas");
        return DFB_ESTATE;
    }
    DFB_CUDA(cudaMemcpyAsync(out_dev, r->d_area.p, (size_t)r->R * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    return DFB_OK;
}

int dfb_calc_pval_dev(dfb_ctx* ctx, const double* areas_dev, int64_t R, double* nullareas_dev, int64_t M, int weight,
                      double* pvalues_dev) {
    DFB_REQUIRE(ctx && (areas_dev || R == 0) && (nullareas_dev || M == 0) && (pvalues_dev || R == 0),
                "dfb_calc_pval_dev: NULL argument");
    DFB_REQUIRE(weight >= 1, "weight must be >= 1");
    return calc_pval_device(ctx, areas_dev, R, nullareas_dev, M, weight, pvalues_dev);
}

int dfb_calc_pval(dfb_ctx* ctx, const double* areas, int64_t R, const double* nullareas, int64_t M, int weight,
                  double* pvalues) {
    DFB_REQUIRE(ctx && (areas || R == 0) && (nullareas || M == 0) && (pvalues || R == 0), "dfb_calc_pval: NULL argument");
    DFB_REQUIRE(weight >= 1, "weight must be >= 1");
    if (R == 0) return DFB_OK;
    DevBuf<double> a, nu, pv;
    DFB_TRY(a.alloc(ctx, (size_t)R));
    DFB_TRY(nu.alloc(ctx, (size_t)std::max<int64_t>(M, 1)));
    DFB_TRY(pv.alloc(ctx, (size_t)R));
    DFB_TRY(h2d(ctx, a.p, areas, (size_t)R * 8));
    DFB_TRY(h2d(ctx, nu.p, nullareas, (size_t)M * 8));
    DFB_TRY(calc_pval_device(ctx, a.p, R, nu.p, M, weight, pv.p));
    return d2h(ctx, pvalues, pv.p, (size_t)R * 8);
}

int dfb_perm_max_dev(dfb_ctx* ctx, const double* nullareas_dev, const int32_t* perm_dev, int64_t M, int64_t B,
                     double fill, double* max_dev) {
    DFB_REQUIRE(ctx && max_dev, "dfb_perm_max_dev: NULL argument");
    return perm_max_device(ctx, nullareas_dev, perm_dev, M, B, fill, max_dev);
}

int dfb_calc_fwer(dfb_ctx* ctx, double cutoff_used, const double* nullareas, const int32_t* permutation, int64_t M,
                  int64_t B, const double* areas, int64_t R, double* fwer) {
    DFB_REQUIRE(ctx && (areas || R == 0) && (fwer || R == 0), "dfb_calc_fwer: NULL argument");
    DFB_REQUIRE(B >= 1, "nPermute must be >= 1");
    if (R == 0) return DFB_OK;
    DevBuf<double> a, nu, mx, fw;
    DevBuf<int32_t> pm;
    DFB_TRY(a.alloc(ctx, (size_t)R));
    DFB_TRY(fw.alloc(ctx, (size_t)R));
    DFB_TRY(nu.alloc(ctx, (size_t)std::max<int64_t>(M, 1)));
    DFB_TRY(pm.alloc(ctx, (size_t)std::max<int64_t>(M, 1)));
    DFB_TRY(mx.alloc(ctx, (size_t)B));
    DFB_TRY(h2d(ctx, a.p, areas, (size_t)R * 8));
    DFB_TRY(h2d(ctx, nu.p, nullareas, (size_t)M * 8));
    DFB_TRY(h2d(ctx, pm.p, permutation, (size_t)M * 4));
    DFB_TRY(perm_max_device(ctx, nu.p, pm.p, M, B, cutoff_used, mx.p));
    {
        KTimer t(ctx, DFB_K_PVAL);
        fwer_kernel<<<(unsigned)ceil_div(R, 128), 128, 0, ctx->stream>>>(a.p, R, mx.p, B, fw.p);
    }
    DFB_CUDA(cudaGetLastError());
    return d2h(ctx, fwer, fw.p, (size_t)R * 8);
}

int dfb_quantile(dfb_ctx* ctx, const dfb_cov* cov, const double* x, int64_t N, double prob, double* out) {
    DFB_REQUIRE(ctx && out, "dfb_quantile: NULL argument");
    if (x) {
        DevBuf<double> tmp;
        DFB_TRY(tmp.alloc(ctx, (size_t)std::max<int64_t>(N, 1)));
        DFB_TRY(h2d(ctx, tmp.p, x, (size_t)N * 8));
        int rc = quantile_device(ctx, tmp.p, N, prob, out);
        cudaStreamSynchronize(ctx->stream);
        return rc;
    }
    DFB_REQUIRE(cov, "dfb_quantile: neither x nor a coverage handle given");
    if (!cov->has_f) {
        set_error("dfb_quantile: no F-statistics cached (run dfb_fstats first)");
        return DFB_ESTATE;
    }
    return quantile_device(ctx, cov->fstats.p, cov->N, prob, out);
}

// Validation hook for the tensor-core screen: runs the dense FP64 kernel and the screen + refine
// path on the same permutations and reports exceedance counts, the candidate count and the number
// of mask bits that differ (must be 0).
int dfb_debug_screen(dfb_ctx* ctx, dfb_cov* cov, const double* mod, int p, const double* mod0, int p0, double adjust_f,
                     const int32_t* perm_idx, int64_t B, double cutoff, int64_t* n_exact_dense,
                     int64_t* n_exact_screen, int64_t* n_candidates, int64_t* n_mask_diff) {
    DFB_REQUIRE(ctx && cov && perm_idx && B > 0, "dfb_debug_screen: bad argument");
    DFB_REQUIRE(cov->y.p != nullptr || cov->y_lazy, "coverageProcessed is missing: run dfb_preprocess first");
    ModelBasis mb;
    DFB_TRY(build_model_basis(mod, cov->n, p, mod0, p0, &mb));
    if (!screen_supported(ctx, cov, mb, -cutoff, cutoff, adjust_f)) {
        set_error("the tensor-core screen does not support this configuration");
        return DFB_ESTATE;
    }
    const int n = cov->n;
    DevBuf<double> qd, q_d;
    DevBuf<int32_t> idx_d;
    DFB_TRY(qd.alloc(ctx, mb.qpad.size()));
    DFB_TRY(h2d(ctx, qd.p, mb.qpad.data(), mb.qpad.size() * sizeof(double)));
    DFB_TRY(q_d.alloc(ctx, mb.q.size()));
    DFB_TRY(h2d(ctx, q_d.p, mb.q.data(), mb.q.size() * sizeof(double)));
    DFB_TRY(idx_d.alloc(ctx, (size_t)B * n));
    DFB_TRY(h2d(ctx, idx_d.p, perm_idx, (size_t)B * n * sizeof(int32_t)));
    PermBatch dense, scr;
    // value capacity: everything for small problems, else 2^27 values (an overflow re-runs with the exact need)
    const size_t hint = std::min<size_t>((size_t)B * (size_t)cov->N, (size_t)1 << 27);
    DFB_TRY(fstat_perm_batch(ctx, cov, mb, qd.p, adjust_f, idx_d.p, B, -cutoff, cutoff, false, hint, &dense));
    int64_t ncand = 0;
    DFB_TRY(fstat_perm_batch_screen(ctx, cov, mb, q_d.p, adjust_f, idx_d.p, B, cutoff, hint, &scr, &ncand));
    int64_t diff = 0;
    DFB_TRY(mask_diff(ctx, dense.mask.p, scr.mask.p, dense.rows * dense.W, &diff));
    if (fused_supported(ctx, cov, mb, -cutoff, cutoff, adjust_f)) {  // the fused kernel must agree as well
        PermBatch fz;
        DFB_TRY(fstat_perm_batch_fused(ctx, cov, mb, q_d.p, adjust_f, idx_d.p, B, cutoff, hint, &fz));
        int64_t diff2 = 0;
        DFB_TRY(mask_diff(ctx, dense.mask.p, fz.mask.p, dense.rows * dense.W, &diff2));
        diff += diff2;
    }
    if (null2_supported(ctx, cov, mb, -cutoff, cutoff, adjust_f)) {  // and the second-generation scan
        PermBatch nz;
        DFB_TRY(fstat_perm_batch_null2(ctx, cov, mb, q_d.p, adjust_f, idx_d.p, B, cutoff, hint, &nz));
        int64_t diff3 = 0;
        DFB_TRY(mask_diff(ctx, dense.mask.p, nz.mask.p, dense.rows * dense.W, &diff3));
        diff += diff3;
    }
    if (n_exact_dense) *n_exact_dense = dense.nvals;
    if (n_exact_screen) *n_exact_screen = scr.nvals;
    if (n_candidates) *n_candidates = ncand;
    if (n_mask_diff) *n_mask_diff = diff;
    return DFB_OK;
}

int dfb_calculate_pvalues(dfb_ctx* ctx, dfb_cov* cov, const double* fstats, const double* mod, int p,
                          const double* mod0, int p0, double adjust_f, const int32_t* perm_idx, int64_t B,
                          double cutoff_lo, double cutoff_hi, int32_t max_region_gap, int32_t max_cluster_gap,
                          dfb_regions** regions, dfb_nulls** nulls, double* pvalues) {
    DFB_REQUIRE(ctx && cov && regions && nulls, "dfb_calculate_pvalues: NULL argument");
    arena_reset(ctx);
    *regions = nullptr;
    *nulls = nullptr;
    DFB_REQUIRE(cutoff_lo <= cutoff_hi, "cutoff pair must be sorted (L <= U)");
    if (fstats) {
        DFB_TRY(cov->fstats.alloc(ctx, (size_t)std::max<int64_t>(cov->N, 1)));
        DFB_TRY(h2d_big(ctx, cov->fstats.p, fstats, (size_t)cov->N * 8));
        cov->has_f = true;
    } else if (!cov->has_f) {
        set_error("dfb_calculate_pvalues: no F-statistics given and none cached");
        return DFB_ESTATE;
    }
    dfb_regions* r = nullptr;
    // The observed-region chain (threshold, segmentation, clusters: dozens of tiny launches and one host
    // synchronisation) runs on the auxiliary stream WHILE the permutation scan occupies the main one:
    // it is enqueued right after the first null-scan launch.
    DFB_TRY(ensure_aux_stream(ctx));
    DFB_CUDA(cudaEventRecord(ctx->ev_main, ctx->stream));  // F-statistics are complete past this point
    int rc_obs = DFB_OK;
    bool obs_done = false;
    const int32_t* h_order = nullptr;
    const std::function<void()> observed = [&]() {
        obs_done = true;
        cudaStream_t main_stream = ctx->stream;
        if (cudaStreamWaitEvent(ctx->aux, ctx->ev_main, 0) != cudaSuccess) {
            set_error("cudaStreamWaitEvent failed");
            rc_obs = DFB_ECUDA;
            return;
        }
        ctx->stream = ctx->aux;
        rc_obs = find_regions_dense(ctx, cov->fstats.p, cov->pos, cutoff_lo, cutoff_hi, max_region_gap, max_cluster_gap,
                                    false, &r, /*defer=*/true);
        if (rc_obs == DFB_OK) {
            // the stable decreasing-area order of the observed regions (:368) does not need the nulls either
            const int64_t R0 = r->R;
            int rc2 = r->d_area_desc.alloc(ctx, (size_t)R0);
            if (rc2 >= 0) rc2 = r->d_order.alloc(ctx, (size_t)R0);
            if (rc2 >= 0) rc2 = order_desc(ctx, r->d_area.p, R0, r->d_order.p, r->d_area_desc.p);
            if (rc2 >= 0) {
                h_order = d2h_async_pinned(ctx, r->d_order.p, (size_t)R0);
                if (!h_order) rc2 = DFB_ENOMEM;
            }
            if (rc2 < 0) rc_obs = rc2;
        }
        if (rc_obs == DFB_OK && cudaEventRecord(ctx->ev_aux, ctx->aux) != cudaSuccess) rc_obs = DFB_ECUDA;
        ctx->stream = main_stream;
    };
    dfb_nulls* nl = nullptr;
    int rc = perm_nulls_impl(ctx, cov, mod, p, mod0, p0, adjust_f, perm_idx, B, cutoff_lo, cutoff_hi, max_region_gap, &nl,
                             &observed);
    if (rc >= 0 && !obs_done) observed();  // no permutation launch happened (B == 0 or nothing kept)
    if (rc >= 0 && rc_obs != DFB_OK) {     // DFB_ENOREGIONS: list(regions = NULL, ...) (:245-251), or an error
        delete nl;
        delete r;
        return rc_obs;
    }
    if (rc >= 0) {
        rc = finalize_regions(ctx, r);  // the observed-region tables arrived while the null scan ran
        if (rc >= 0 && cudaStreamWaitEvent(ctx->stream, ctx->ev_aux, 0) != cudaSuccess) rc = DFB_ECUDA;
        if (rc < 0) delete nl;
    }
    if (rc < 0) {
        delete r;
        return rc;
    }
    // order by decreasing area (stable), then p-values against the (duplicated) null areas; the
    // region areas are still on the device, order and p-values come back with one synchronisation
    const int64_t R = r->R;
    DevBuf<double> p_d;
    auto fail = [&](int code) {
        delete r;
        delete nl;
        return code;
    };
    if ((rc = p_d.alloc(ctx, (size_t)R)) < 0) return fail(rc);
    const double* h_p = nullptr;
    if (nl->count > 0) {  // no nulls at all -> NA (:408-419)
        // the regions are sorted already: rank the nulls against them instead of sorting the nulls
        if ((rc = pval_by_hist(ctx, r->d_area_desc.p, R, nl->area.p, nl->count, 2, p_d.p)) < 0) return fail(rc);
        h_p = d2h_async_pinned(ctx, p_d.p, (size_t)R);
    }
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess || !h_order || (nl->count > 0 && !h_p)) {
        set_error("dfb_calculate_pvalues: read-back failed");
        return fail(DFB_ECUDA);
    }
    r->order.assign(h_order, h_order + R);  // arrived on the auxiliary stream (finalize_regions synchronised it)
    if (h_p) r->pvalues.assign(h_p, h_p + R);
    else r->pvalues.assign((size_t)R, nan(""));
    if (pvalues) memcpy(pvalues, r->pvalues.data(), (size_t)R * 8);
    *regions = r;
    *nulls = nl;
    return DFB_OK;
}

}  // extern "C"
