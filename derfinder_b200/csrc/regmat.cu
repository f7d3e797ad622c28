// K6 -- regionMatrix: per region x sample coverage sums.
//
// Reference: .regionMatrixByChr (R/regionMatrix.R:252-270) with getRegionCoverage
// (R/getRegionCoverage.R:135-160): rows indexStart..indexEnd of the filtered coverage are
// gathered per region and summed per sample (`colSums`), then divided by L.  In R this goes
// through one data.frame per region; here one warp owns one (region, sample) pair and streams the
// region's slice of the sample-major matrix (coalesced 128 B lines, each byte read once).
// Integer coverage is accumulated exactly in int64 (any order is bit-exact); double coverage
// (library-size normalised) uses a fixed-shape warp tree -- R's colSums accumulates in long
// double, so that case is checked to 1e-12 relative instead of bit for bit.
#include <math.h>

#include <algorithm>

#include "data.cuh"

namespace dfb {

template <typename V>
__global__ void __launch_bounds__(256) region_sums_kernel(const V* __restrict__ x, int64_t ld, int n,
                                                          const int32_t* __restrict__ is, const int32_t* __restrict__ ie,
                                                          const int32_t* __restrict__ order, int64_t R,
                                                          const double* __restrict__ Lvec, int n_l,
                                                          int every_column, double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warps_per_grid = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int s = blockIdx.y;
    const V* __restrict__ col = x + (size_t)s * ld;
    for (int64_t r = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); r < R; r += warps_per_grid) {
        const int64_t src = order ? order[r] : r;
        const int64_t a = (int64_t)is[src] - 1, e = ie[src];
        double total;
        if (sizeof(V) == 4) {
            long long acc = 0;
            for (int64_t i = a + lane; i < e; i += 32) acc += (long long)col[i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            total = (double)acc;
        } else {
            double acc = 0.0;
            for (int64_t i = a + lane; i < e; i += 32) acc += (double)col[i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            total = acc;
        }
        if (lane == 0) {
            // covMat / L for scalar L; a vector L only divides the LAST column: `for (i in
            // length(L))` (R/regionMatrix.R:267)
            // (railMatrix divides every sample's column by its own L, R/railMatrix.R:289)
            if (n_l == 1) total = __ddiv_rn(total, Lvec[0]);
            else if (n_l == n && (every_column || s == n - 1)) total = __ddiv_rn(total, Lvec[s]);
            out[(size_t)s * R + r] = total;
        }
    }
}

// sum(Views(x_s, IRanges(start, end))) / L_s for double coverage, in the order IRanges' viewSums
// walks an Rle: one product value * (run length inside the view) per run of identical adjacent
// values, accumulated left to right (.railMatChrRegionCov, R/railMatrix.R:285-289, after
// `regCov / mappedPerXM` made the values non-integer).  One warp per (view, sample): every chunk of 32
// bases is loaded coalesced, run heads are found with one ballot, the per-run products are formed in
// parallel by the head lanes and only the final chain of additions (one per run) is sequential.
__global__ void __launch_bounds__(256) view_sums_kernel(const double* __restrict__ x, int64_t ld, int n,
                                                        const int32_t* __restrict__ vs, const int32_t* __restrict__ ve,
                                                        int64_t R, const double* __restrict__ Lvec, int n_l,
                                                        double* __restrict__ out) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int64_t warps_per_grid = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int s = blockIdx.y;
    const double* __restrict__ col = x + (size_t)s * ld;
    for (int64_t r = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); r < R; r += warps_per_grid) {
        const int64_t a = (int64_t)vs[r] - 1, e = ve[r];
        double acc = 0.0, open_v = 0.0;  // the run still open at the end of the previous chunk
        int64_t open_len = 0;
        for (int64_t base = a; base < e; base += 32) {
            const int64_t i = base + lane;
            const bool valid = i < e;
            const double v = valid ? col[i] : 0.0;
            double pv = __shfl_up_sync(full, v, 1);
            if (lane == 0) pv = open_v;
            const bool head = valid && (i == a || v != pv);
            const unsigned hm = __ballot_sync(full, head);
            const int cnt = (int)min((int64_t)32, e - base);
            if (hm == 0) {
                open_len += cnt;
                continue;
            }
            open_len += __ffs(hm) - 1;
            if (open_len > 0) acc = __dadd_rn(acc, __dmul_rn(open_v, (double)open_len));
            const unsigned above = hm & ~((2u << lane) - 1u);
            const int next_head = above ? __ffs(above) - 1 : cnt;
            const double prod = __dmul_rn(v, (double)(next_head - lane));
            const int last = 31 - __clz(hm);
            unsigned m = hm & ~(1u << last);
            while (m) {
                const int b = __ffs(m) - 1;
                m &= m - 1;
                acc = __dadd_rn(acc, __shfl_sync(full, prod, b));
            }
            open_v = __shfl_sync(full, v, last);
            open_len = cnt - last;
        }
        if (open_len > 0) acc = __dadd_rn(acc, __dmul_rn(open_v, (double)open_len));
        if (lane == 0) {
            if (n_l == 1) acc = __ddiv_rn(acc, Lvec[0]);
            else if (n_l == n) acc = __ddiv_rn(acc, Lvec[s]);
            out[(size_t)s * R + r] = acc;
        }
    }
}

static int region_matrix_impl(dfb_ctx* ctx, const dfb_cov* cov, const int32_t* is_dev, const int32_t* ie_dev,
                              const int32_t* order_dev, int64_t R, const double* L, int n_l, double* out,
                              bool views = false) {
    DFB_REQUIRE(n_l == 0 || n_l == 1 || n_l == cov->n,
                "Invalid 'L' value: it has to be of length 1 or equal to the number of samples");
    if (R == 0) return DFB_OK;
    DevBuf<double> o, Ld;
    DFB_TRY(o.alloc(ctx, (size_t)R * cov->n));
    if (n_l > 0) {
        DFB_TRY(Ld.alloc(ctx, (size_t)n_l));
        DFB_TRY(h2d(ctx, Ld.p, L, (size_t)n_l * sizeof(double)));
    }
    dim3 grid((unsigned)std::min<int64_t>(ceil_div(R, 8), (int64_t)ctx->sm_count * 8), (unsigned)cov->n);
    {
        KTimer t(ctx, DFB_K_REGMAT);
        if (cov->dtype == DFB_I32)  // exact in int64: identical to any double summation order below 2^53
            region_sums_kernel<int32_t><<<grid, 256, 0, ctx->stream>>>(cov->xi.p, cov->ld, cov->n, is_dev, ie_dev,
                                                                       order_dev, R, Ld.p, n_l, views ? 1 : 0, o.p);
        else if (views)
            view_sums_kernel<<<grid, 256, 0, ctx->stream>>>(cov->xd.p, cov->ld, cov->n, is_dev, ie_dev, R, Ld.p, n_l,
                                                            o.p);
        else
            region_sums_kernel<double><<<grid, 256, 0, ctx->stream>>>(cov->xd.p, cov->ld, cov->n, is_dev, ie_dev,
                                                                      order_dev, R, Ld.p, n_l, 0, o.p);
    }
    DFB_CUDA(cudaGetLastError());
    return d2h(ctx, out, o.p, (size_t)R * cov->n * sizeof(double));
}

}  // namespace dfb

using namespace dfb;

extern "C" {

int dfb_region_matrix(dfb_ctx* ctx, const dfb_cov* cov, const int32_t* index_start, const int32_t* index_end,
                      int64_t R, const double* L, int n_l, double* out) {
    DFB_REQUIRE(ctx && cov && (out || R == 0), "dfb_region_matrix: NULL argument");
    DFB_REQUIRE(R == 0 || (index_start && index_end), "dfb_region_matrix: NULL index ranges");
    for (int64_t r = 0; r < R; ++r)
        DFB_REQUIRE(index_start[r] >= 1 && index_end[r] >= index_start[r] && index_end[r] <= cov->N,
                    "region %lld: index range [%d, %d] outside 1..%lld", (long long)r + 1, index_start[r],
                    index_end[r], (long long)cov->N);
    DevBuf<int32_t> a, b;
    DFB_TRY(a.alloc(ctx, (size_t)std::max<int64_t>(R, 1)));
    DFB_TRY(b.alloc(ctx, (size_t)std::max<int64_t>(R, 1)));
    DFB_TRY(h2d(ctx, a.p, index_start, (size_t)R * 4));
    DFB_TRY(h2d(ctx, b.p, index_end, (size_t)R * 4));
    return region_matrix_impl(ctx, cov, a.p, b.p, nullptr, R, L, n_l, out);
}

int dfb_view_sums(dfb_ctx* ctx, const dfb_cov* cov, const int32_t* start, const int32_t* end, int64_t R,
                  const double* L, int n_l, double* out) {
    DFB_REQUIRE(ctx && cov && (out || R == 0), "dfb_view_sums: NULL argument");
    DFB_REQUIRE(R == 0 || (start && end), "dfb_view_sums: NULL ranges");
    for (int64_t r = 0; r < R; ++r)
        DFB_REQUIRE(start[r] >= 1 && end[r] >= start[r] && end[r] <= cov->N,
                    "view %lld: range [%d, %d] outside 1..%lld", (long long)r + 1, start[r], end[r],
                    (long long)cov->N);
    DevBuf<int32_t> a, b;
    DFB_TRY(a.alloc(ctx, (size_t)std::max<int64_t>(R, 1)));
    DFB_TRY(b.alloc(ctx, (size_t)std::max<int64_t>(R, 1)));
    DFB_TRY(h2d(ctx, a.p, start, (size_t)R * 4));
    DFB_TRY(h2d(ctx, b.p, end, (size_t)R * 4));
    return region_matrix_impl(ctx, cov, a.p, b.p, nullptr, R, L, n_l, out, true);
}

int dfb_region_matrix_r(dfb_ctx* ctx, const dfb_cov* cov, const dfb_regions* r, const double* L, int n_l,
                        double* out) {
    DFB_REQUIRE(ctx && cov && r && (out || r->R == 0), "dfb_region_matrix_r: NULL argument");
    DevBuf<int32_t> ord;
    if (!r->order.empty()) {
        DFB_TRY(ord.alloc(ctx, r->order.size()));
        DFB_TRY(h2d(ctx, ord.p, r->order.data(), r->order.size() * 4));
    }
    return region_matrix_impl(ctx, cov, r->d_is.p, r->d_ie.p, r->order.empty() ? nullptr : ord.p, r->R, L, n_l, out);
}

}  // extern "C"
