// Device-resident objects behind the opaque C handles.
#pragma once

#include "common.cuh"

// Position index of one chromosome: logical Rle over the chromosome, kept as host run lengths
// (metadata, O(#runs)) plus two small device arrays describing the TRUE runs:
//   run_gstart[k]  genomic 1-based start of TRUE run k
//   run_cum[k]     number of kept bases before TRUE run k  (run_cum[K] = N)
// `which(position)` (R/findRegions.R:221, 275 MB for 72 M bases in R) is never materialised:
// index -> genomic coordinate is a binary search in run_cum.
struct PositionIndex {
    std::vector<int32_t> len;  // normalised run lengths (alternating values)
    int first = 0;             // value of the first run
    int64_t chr_len = 0;
    int64_t N = 0;  // sum of TRUE run lengths
    int64_t K = 0;  // number of TRUE runs
    std::vector<int32_t> h_gstart;
    std::vector<int64_t> h_cum;
    dfb::DevBuf<int32_t> run_gstart;
    dfb::DevBuf<int64_t> run_cum;

    int build(dfb_ctx* ctx, const int32_t* pos_len, int64_t nruns, int pos_first);
};

struct dfb_cov {
    dfb_ctx* ctx = nullptr;
    int n = 0;       // samples
    int64_t N = 0;   // kept bases
    int64_t ld = 0;  // leading dimension of the sample-major matrices (>= N, multiple of 64)
    int dtype = DFB_I32;
    dfb::DevBuf<int32_t> xi;  // [n][ld] raw filtered coverage (dtype I32)
    dfb::DevBuf<double> xd;   // [n][ld] raw filtered coverage (dtype F64, e.g. normalised)
    dfb::DevBuf<double> y;    // [n][ld] log2(x + scalefac)   (after dfb_preprocess; see y_lazy)
    // Integer coverage: dfb_preprocess does not write the log2 matrix at all (8 bytes per element that the hot
    // path would write once and read once).  The observed-F kernel and the refinement of the null scan read the
    // int32 coverage through the host-built log2 table instead; ensure_y() materialises `y` for everything else
    // (coverageProcessed read-back, the dense FP64 fallback, the first-generation tensor kernels).
    bool y_lazy = false;
    dfb::DevBuf<double> ym;   // [N] (sum_k y[k]) / n in sample order: the row mean of the F-stat contract
    bool has_ym = false;
    dfb::DevBuf<double> mean; // [N] meanCoverage
    bool has_mean = false;
    int G = 0;
    dfb::DevBuf<double> gmean;  // [G][ld]
    dfb::DevBuf<double> fstats; // [N] cached by dfb_fstats
    bool has_f = false;
    double scalefac = 32;
    PositionIndex pos;
    // operand cache of the K2 null scan (fstat_tc2.cu): centred log2 coverage as fp16 rows [a16_rows][a16_ka],
    // |yc|^2 and the row mean in FP64; valid while `y` is unchanged (a16_center: -1 = stale)
    dfb::DevBuf<uint16_t> a16;
    dfb::DevBuf<double> a16_yy, a16_m;
    int a16_center = -1, a16_ka = 0;
    int64_t a16_rows = 0;
};

struct dfb_regions {
    dfb_ctx* ctx = nullptr;
    int64_t R = 0, R_up = 0;
    bool basic = false;
    // host copies (R is small: O(#regions))
    std::vector<int32_t> start, end, index_start, index_end, cluster;
    std::vector<double> value, area, cluster_l, pvalues;
    std::vector<int32_t> order;
    // device copies of the index ranges (kept for view means / region matrix) and of the areas
    dfb::DevBuf<int32_t> d_is, d_ie;
    dfb::DevBuf<double> d_area;
    // asynchronous read-backs not yet copied into the vectors above (pinned arena memory)
    struct Pending {
        const int32_t *gs = nullptr, *ge = nullptr, *cf = nullptr;
        const double* cl = nullptr;
        const double *val = nullptr, *area = nullptr;
        const int32_t *is = nullptr, *ie = nullptr;
        bool active = false;
        cudaStream_t stream = nullptr;  // stream the read-backs were enqueued on
    } pend;
    dfb::DevBuf<double> d_area_desc;  // areas in decreasing order (after dfb_calculate_pvalues)
    dfb::DevBuf<int32_t> d_order;     // d_area_desc[i] = d_area[d_order[i]]
};

struct dfb_nulls {
    dfb_ctx* ctx = nullptr;
    int64_t B = 0;
    int64_t count = 0;
    dfb::DevBuf<double> area, stat;  // [count]
    dfb::DevBuf<double> area_sorted; // [count] ascending (kept from the p-value ranking; pooled across GPUs)
    dfb::DevBuf<double> area_merge;  // [count] abs(stat) * width (R/mergeResults.R:231), built on first use
    dfb::DevBuf<int32_t> width, perm;  // [count], perm ids 1-based
    std::vector<int64_t> per_perm;   // [B]
};

namespace dfb {

// ---- cov.cu
// the log2 matrix of a handle whose dfb_preprocess left it un-materialised (dfb_cov::y_lazy); no-op otherwise
int ensure_y(dfb_ctx* ctx, const dfb_cov* cov);
// the context's host-built log2 table for this scalefac (64 Ki entries), rebuilt when the scalefac changed
int ensure_lut(dfb_ctx* ctx, double scalefac);
constexpr int LUT_SMEM = 4096;  // table entries the streaming kernels keep in shared memory

// ---- fstat_tc2.cu
// Which tensor-core null scan: the second generation (one fp16 product per chunk, TMA-fed) wins once the
// contraction is long enough to matter; with up to 16 samples (K = 16: a single MMA per chunk) the scan is bound by
// the FP64 refinement of frequent candidates and the first fused kernel, which keeps the FP64 tile in shared
// memory, is faster (C2: 1.37 ms against 2.81 ms, gpurun_out/r2_am_sweep_c2.txt).
inline bool null2_wanted(const dfb_ctx* ctx, int n) { return ctx->opt_null2 == 2 || (ctx->opt_null2 == 1 && n > 16); }
struct PermBatch;
// zero the flagged groups of a batch that used the context's kept-zero mask (no-op otherwise)
int null2_release_mask(dfb_ctx* ctx, PermBatch* batch);

// ---- fstat.cu
struct ModelBasis {
    int n = 0, p = 0, p0 = 0, pc = 0, chunks = 0, pp = 0;  // pp = padded columns = chunks * pc
    bool centered = false;
    std::vector<double> q;     // [n][p] row-major (un-padded)
    std::vector<double> qpad;  // [n][pp] row-major, zero padded
};
int build_model_basis(const double* mod, int n, int p, const double* mod0, int p0, ModelBasis* out);

// Dense F for the identity labelling (observed statistics): F[N] on the device.
int fstat_observed(dfb_ctx* ctx, const dfb_cov* cov, const ModelBasis& mb, double adjust_f, double* f_dev);

// Result of one batch of permutations in "mask + compacted exceedance values" form.
struct PermBatch {
    int64_t rows = 0;      // number of mask rows (= permutations in the batch, x2 when two-sided)
    int64_t W = 0;         // words per row
    int64_t tiles = 0;     // value tiles per row
    int tile_bases = 0;    // bases per value tile
    DevBuf<uint32_t> mask; // [rows][W]
    DevBuf<int64_t> off;   // [rows][tiles] offset of the tile's first exceedance value
    DevBuf<double> vals;   // compacted F values
    DevBuf<unsigned char> gflag;  // optional [rows][ceil(W / 32)]: 1 when the 32-word group holds a non-empty word
    bool mask_cached = false;     // mask / gflag are the context's kept-zero buffers: null2_release_mask() after use
    int64_t nvals = 0;
    // deferred overflow check (fused path): pinned copy of the value cursor, valid after the next
    // stream synchronisation; the batch is good iff *used_host <= capacity
    const unsigned long long* used_host = nullptr;
    size_t capacity = 0;
};

int fstat_tile_bases(int n);
// One batch of permutations -> mask + compacted exceedances.  idx_dev: [B][n] on the device.
int fstat_perm_batch(dfb_ctx* ctx, const dfb_cov* cov, const ModelBasis& mb, const double* qpad_dev,
                     double adjust_f, const int32_t* idx_dev, int64_t B, double lo, double hi, bool two_sided,
                     size_t capacity_hint, PermBatch* out);

// ---- fstat_tc.cu: tcgen05 screening contraction + exact FP64 refinement of the candidates
bool screen_supported(const dfb_ctx* ctx, const dfb_cov* cov, const ModelBasis& mb, double lo, double hi,
                      double adjust_f);
int fstat_perm_batch_screen(dfb_ctx* ctx, const dfb_cov* cov, const ModelBasis& mb, const double* q_dev /*[n][p]*/,
                            double adjust_f, const int32_t* idx_dev, int64_t B, double hi, size_t capacity_hint,
                            PermBatch* out, int64_t* ncand_out);

bool fused_supported(const dfb_ctx* ctx, const dfb_cov* cov, const ModelBasis& mb, double lo, double hi,
                     double adjust_f);
int fstat_perm_batch_fused(dfb_ctx* ctx, const dfb_cov* cov, const ModelBasis& mb, const double* q_dev /*[n][p]*/,
                           double adjust_f, const int32_t* idx_dev, int64_t B, double hi, size_t capacity_hint,
                           PermBatch* out, bool defer_check = false);

// ---- fstat_tc2.cu: second-generation null scan (TMA-fed fp16 screen, permutation operand resident in smem)
bool null2_supported(const dfb_ctx* ctx, const dfb_cov* cov, const ModelBasis& mb, double lo, double hi,
                     double adjust_f);
int fstat_perm_batch_null2(dfb_ctx* ctx, const dfb_cov* cov, const ModelBasis& mb, const double* q_dev /*[n][p]*/,
                           double adjust_f, const int32_t* idx_dev, int64_t B, double hi, size_t capacity_hint,
                           PermBatch* out, bool defer_check = false);

int mask_diff(dfb_ctx* ctx, const uint32_t* a, const uint32_t* b, int64_t words, int64_t* diff);

// ---- regions.cu
int exclusive_scan_i64(dfb_ctx* ctx, const int64_t* in, int64_t* out, int64_t n, int64_t* total_dev, int fam);
// Regions of `rows` mask rows: device outputs sized by `total`; row_counts (host) = regions per row.
struct RegionSet {
    int64_t total = 0;
    std::vector<int64_t> row_counts;
    DevBuf<double> area, stat;
    DevBuf<int32_t> width, istart, iend, perm;
};
int find_regions_rows(dfb_ctx* ctx, int64_t N, int64_t rows, int sides, const uint32_t* mask, const uint32_t* seg,
                      const double* dense, int64_t dense_stride, const PermBatch* compact, bool want_index,
                      bool want_perm, RegionSet* out);
int find_regions_dense(dfb_ctx* ctx, const double* x_dev, const PositionIndex& pos, double lo, double hi,
                       int32_t max_region_gap, int32_t max_cluster_gap, bool basic, dfb_regions** out,
                       bool defer = false);
int finalize_regions(dfb_ctx* ctx, dfb_regions* r);
int build_segment_starts(dfb_ctx* ctx, const PositionIndex& pos, int32_t max_region_gap, DevBuf<uint32_t>* seg);

}  // namespace dfb
