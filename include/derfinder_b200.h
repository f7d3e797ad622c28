/*
 * derfinder_b200 -- C ABI of the B200-native engine for derfinder's single-base F-statistic
 * pipeline and expressed-region (ER) path.
 *
 * This header is the drop-in boundary.  The reference (lcolladotor/derfinder v1.35.0) is pure R
 * and has no FFI of its own; its seam is (i) the exported R functions and (ii) the call
 * `fstats.apply(index, data, mod, mod0, lowMemDir, scalefac, method, adjustF)` dispatched by
 * `bplapply` (R/calculateStats.R:123-127, R/calculatePvalues.R:325-329).  Each entry point below
 * names the reference interface it replaces; `INTEGRATION.md` shows the `.Call` stubs and the R
 * glue that bind them under the unchanged R signatures.
 *
 * Conventions
 *   - plain C: pointers + sizes, no R / torch / C++ types.  Every function returns 0 on success
 *     or a negative DFB_E* code; `dfb_last_error()` returns the message (thread-local).
 *   - all inputs are caller-owned and read-only; outputs go to caller-allocated buffers
 *     ("two-call sizing": query the count, allocate, fetch).
 *   - indices and coordinates are 1-based where the reference's are (indexStart, indexEnd,
 *     start, end, cluster ids, permutation ids); permutation row indices `perm_idx` are 0-based
 *     (`sample(n) - 1`).
 *   - coverage matrices are sample-major: column s (one sample, all bases) is contiguous, like
 *     the columns of the reference's DataFrame.
 *   - "host" pointers are ordinary memory; functions ending in `_dev` take CUDA device pointers
 *     on the context's device and enqueue on the context's stream without synchronising.
 *   - there is no CPU fallback: without a CUDA device `dfb_create` fails.
 */
#ifndef DERFINDER_B200_H
#define DERFINDER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DFB_VERSION 100

/* error codes */
#define DFB_OK 0
#define DFB_EINVAL (-1)   /* bad argument (the R glue's stopifnot()/stop() cases) */
#define DFB_ECUDA (-2)    /* CUDA runtime error */
#define DFB_ENOMEM (-3)   /* host or device allocation failed */
#define DFB_ESTATE (-4)   /* handle is missing a prerequisite (e.g. preprocess not run) */
#define DFB_ENOREGIONS 1  /* not an error: nothing passed the cutoff (reference returns NULL) */

/* element types of coverage values (R integer / R numeric) */
#define DFB_I32 0
#define DFB_F64 1

/* filter kinds, R/filterData.R:89 `filter = "one" | "mean"` */
#define DFB_FILTER_NONE 0 /* cutoff = NULL */
#define DFB_FILTER_ONE 1
#define DFB_FILTER_MEAN 2

/* kernels whose device time the context can report (dfb_kernel_time) */
#define DFB_K_PREP 1       /* K1 filter / expand / log2 */
#define DFB_K_FSTAT 2      /* K2 F-stat scan (FP64 exact kernel) */
#define DFB_K_FSTAT_TC 3   /* K2 tcgen05 screening contraction */
#define DFB_K_REGIONS 4    /* K3/K4 segmentation, areas, clusters */
#define DFB_K_PVAL 5       /* K5 sort + rank */
#define DFB_K_REGMAT 6     /* K6 region x sample sums */
#define DFB_K_COUNT 8

typedef struct dfb_ctx dfb_ctx;         /* one per GPU: stream, scratch, timers            */
typedef struct dfb_cov dfb_cov;         /* device-resident filtered coverage of ONE chromosome
                                           (+ position, means, and after dfb_preprocess the
                                           log2 matrix)                                      */
typedef struct dfb_regions dfb_regions; /* device-resident result of a region search        */
typedef struct dfb_nulls dfb_nulls;     /* device-resident null regions of a permutation run */

/* ---------------------------------------------------------------- context ------------------- */

const char* dfb_last_error(void);
int dfb_version(void);
/* number of visible CUDA devices (0 when there is no driver / GPU). */
int dfb_device_count(void);
/* `stream` is a cudaStream_t (or NULL: the library creates its own non-blocking stream). */
int dfb_create(int device, void* stream, dfb_ctx** out);
void dfb_destroy(dfb_ctx* ctx);
int dfb_sync(dfb_ctx* ctx);
/* Kernel launches issued by this context since creation / the last reset. */
int64_t dfb_launch_count(dfb_ctx* ctx);
void dfb_reset_counters(dfb_ctx* ctx);
/* When enabled every launch of a DFB_K_* kernel family is bracketed by CUDA events on the
 * context's stream; dfb_kernel_time synchronises and returns the summed device ms and the number
 * of launches since the last reset.  on = 1: every family; on > 1: a mask of (1 << DFB_K_*), so a
 * benchmark can time its dominant kernel live without paying two event records on every launch. */
int dfb_enable_timing(dfb_ctx* ctx, int on);
int dfb_kernel_time(dfb_ctx* ctx, int kernel, double* ms, int64_t* launches);
/* Engine knobs (all default to the fastest exact configuration):
 *   "screen"     0/1  use the tcgen05 split-bf16 screening contraction in front of the FP64
 *                     refinement for permutation nulls (default 1 when n fits; results are
 *                     identical either way)
 *   "fused"      0/1  screen and refinement in one persistent kernel (default 1 when the tile
 *                     fits in shared memory; 0 = separate screen and refinement kernels)
 *   "null_capacity_mb"  initial capacity of the exceedance value buffer
 *   "perm_batch"        permutations per launch (0 = auto)
 *   "scratch_gb"        upper bound for per-call device scratch (default 24)
 *   "null2"       0/1/2  second-generation tensor-core null scan: 1 = from 17 samples on (default),
 *                     2 = always, 0 = never (the first fused kernel)
 *   "lazy_y"      0/1  integer coverage: dfb_preprocess does not write the log2 matrix; it is
 *                     produced on demand (default 1; results identical)
 *   "mask_cache"  0/1  mask rows of the null scan kept zero between calls instead of a memset per
 *                     call (default 1)
 *   "filter_events" 0/1  dfb_filter_rle of integer coverage without a divisor in run space
 *                     (default 1; 0 = tile expansion kernel; results identical)
 * Developer knobs of the fused null kernel (geometry overrides for measurements; results are
 * identical for every valid setting, tests/test_gpu_parity.py::test_fused_kernel_variants_agree):
 *   "chunk_perms" permutations per MMA chunk, "b_stages" B-tile stages (1-4), "tmem_stages"
 *   accumulator stages (1-2), "epi_warps" epilogue warps per TMEM quarter (1-4), "ys_smem" FP64
 *   tile resident (1) or streamed (0), "k2_waves" CTAs launched per resident slot (default 16;
 *   1 = fully persistent grid); "dbg" bit flags switch parts of the kernel OFF for timing
 *   experiments (results then invalid; 64 = print clock64 phase counters only). */
int dfb_set_option(dfb_ctx* ctx, const char* key, double value);

/* ---------------------------------------------------------------- coverage in --------------- */

/* filterData() -- R/filterData.R:89-241.
 * Input: n Rle columns (values[s] of `dtype`, lengths[s] int32, nruns[s]) each of total length
 * `len` (= chromosome length, or = sum(prev_index) when a previous logical index is given as
 * the run lengths `prev_len[0..prev_nruns)` whose first run has value `prev_first`).
 * `mapped_per_xm` (NULL or n doubles = totalMapped/targetSize) divides each column first
 * (:119-130); the stored coverage is then double.  filter: DFB_FILTER_*; strict `>` (:139-155).
 * The result holds the kept rows (N x n), the final position index, and meanCoverage. */
int dfb_filter_rle(dfb_ctx* ctx, int n, int dtype, const void* const* values,
                   const int32_t* const* lengths, const int64_t* nruns, int64_t len,
                   const double* mapped_per_xm, int filter, double cutoff,
                   const int32_t* prev_len, int64_t prev_nruns, int prev_first,
                   dfb_cov** out);

/* Already filtered coverage (what loadCoverage(cutoff=) hands to preprocessCoverage,
 * R/preprocessCoverage.R:131-165): dense sample-major host matrix cov[s*ld + i], i < N, and the
 * position index as run lengths over the chromosome. */
int dfb_cov_from_dense(dfb_ctx* ctx, int n, int64_t N, int dtype, const void* cov, int64_t ld,
                       const int32_t* pos_len, int64_t pos_nruns, int pos_first, dfb_cov** out);
/* Same, from n Rle columns of length N (the DataFrame the R glue holds). */
int dfb_cov_from_rle(dfb_ctx* ctx, int n, int dtype, const void* const* values,
                     const int32_t* const* lengths, const int64_t* nruns, int64_t N,
                     const int32_t* pos_len, int64_t pos_nruns, int pos_first, dfb_cov** out);
/* Device-resident variant used by benchmarks: cov_dev is a device pointer, copied D2D. */
int dfb_cov_from_dense_dev(dfb_ctx* ctx, int n, int64_t N, int dtype, const void* cov_dev,
                           int64_t ld, const int32_t* pos_len, int64_t pos_nruns, int pos_first,
                           dfb_cov** out);
/* Zero-copy variant: when cov_dev is laid out like the engine's own matrix (ld a multiple of 64
 * elements, base aligned to 256 bytes) the handle READS IT IN PLACE -- the caller keeps it alive and
 * unmodified until dfb_cov_free; any other layout is copied as above. */
int dfb_cov_view_dense_dev(dfb_ctx* ctx, int n, int64_t N, int dtype, const void* cov_dev,
                           int64_t ld, const int32_t* pos_len, int64_t pos_nruns, int pos_first,
                           dfb_cov** out);
void dfb_cov_free(dfb_cov* cov);

/* Shapes: N kept bases, n samples, chromosome length, dtype of the stored coverage, number of
 * runs of the position index, G group-mean vectors (0 before dfb_preprocess). */
int dfb_cov_info(const dfb_cov* cov, int64_t* N, int* n, int64_t* chr_len, int* dtype,
                 int64_t* pos_nruns, int* n_groups, int* has_processed);
/* position logical Rle: run lengths; *first = value of the first run (runs alternate). */
int dfb_cov_get_position(const dfb_cov* cov, int32_t* lengths, int* first);
/* one filtered coverage column (N values of the stored dtype) */
int dfb_cov_get_column(const dfb_cov* cov, int sample, void* out);
int dfb_cov_get_mean(const dfb_cov* cov, double* out);                 /* meanCoverage [N] */

/* preprocessCoverage() -- R/preprocessCoverage.R:97-260 (the colsubset re-filter is a
 * dfb_filter_rle call made by the glue first).  Computes on the device, in one pass over the
 * filtered coverage: meanCoverage (if absent, :191-193), the G group means (:197-206;
 * group_of_sample[s] in 0..G-1, or NULL/G=0) and Y = log2(x + scalefac) (:209-212).
 * For integer coverage Y is not stored: the F-stat kernels read the coverage through the same
 * host-built log2 table; dfb_cov_get_processed_column (coverageProcessed) produces it on demand. */
int dfb_preprocess(dfb_ctx* ctx, dfb_cov* cov, double scalefac, const int32_t* group_of_sample,
                   int n_groups);
int dfb_cov_get_processed_column(const dfb_cov* cov, int sample, double* out); /* log2 column */
int dfb_cov_get_group_mean(const dfb_cov* cov, int group, double* out);        /* [N] */
/* chunk lengths of `mclapplyIndex` (R/preprocessCoverage.R:171-188, 218-227); chunksize <= 0
 * means NULL -> ceiling(N / mc_cores).  Returns the number of chunks; fills lens if non-NULL. */
int64_t dfb_chunk_lengths(int64_t N, double chunksize, int mc_cores, int64_t* lens, int64_t cap);

/* ---------------------------------------------------------------- F statistics -------------- */

/* calculateStats() -- R/calculateStats.R:70-132, replacing derfinderHelper::fstats.apply.
 * mod [n x p], mod0 [n x p0] column-major doubles (R matrices).  F (N doubles) is written to
 * `fstats_out` (host) if non-NULL and kept on the device inside `cov` for later calls.
 * F = ((rss0-rss1)/(p-p0)) / (adjustF + rss1/(n-p))  (tests/testthat/test_Fstats.R:52-63). */
int dfb_fstats(dfb_ctx* ctx, dfb_cov* cov, const double* mod, int p, const double* mod0, int p0,
               double adjust_f, double* fstats_out);
/* Split form for callers that do not read the statistics right away (the R glue can hand back a
 * lazily materialised vector: calculatePvalues() uses the device copy, R/calculatePvalues.R:234):
 * dfb_fstats_begin returns once the kernel is queued and downloads into fstats_out in the
 * background -- its own stream, pinned stages and host threads -- while the caller goes on (normally
 * to dfb_calculate_pvalues).  fstats_out is complete, and may be touched, after dfb_fstats_wait;
 * dfb_cov_free and any other bulk transfer of the context wait for it as well. */
int dfb_fstats_begin(dfb_ctx* ctx, dfb_cov* cov, const double* mod, int p, const double* mod0,
                     int p0, double adjust_f, double* fstats_out);
int dfb_fstats_wait(dfb_ctx* ctx);
/* Orthonormal nested basis the engine uses for (mod0, mod): q is [n x p] column-major, first p0
 * columns span mod0.  *centered = 1 when the constant vector lies in span(mod0) and rows are
 * mean-centred before the contraction.  Exposed for the parity tests. */
int dfb_model_basis(const double* mod, int n, int p, const double* mod0, int p0, double* q,
                    int* centered);

/* ---------------------------------------------------------------- regions ------------------- */

/* findRegions() -- R/findRegions.R:116-296 (smooth=FALSE) incl. .getSegmentsRle (:363-396) and
 * .clusterMakerRle (:445-475).  `x` are N host doubles (F-stats or meanCoverage) or NULL to use
 * the F-stats cached in `cov` by dfb_fstats; position comes from `cov`.  cutoff_lo/cutoff_hi are
 * the sorted pair (L, U); a scalar cutoff c is passed as (-c, c) (:377-381).
 * basic != 0 returns only area/width/stat rows (:273-278).
 * Returns DFB_ENOREGIONS (and *out = NULL) when nothing passes. */
int dfb_find_regions(dfb_ctx* ctx, const dfb_cov* cov, const double* x, double cutoff_lo,
                     double cutoff_hi, int32_t max_region_gap, int32_t max_cluster_gap, int basic,
                     dfb_regions** out);
/* Position-only variant (no coverage handle): x has N = number of TRUE positions. */
int dfb_find_regions_pos(dfb_ctx* ctx, const double* x, int64_t N, const int32_t* pos_len,
                         int64_t pos_nruns, int pos_first, double cutoff_lo, double cutoff_hi,
                         int32_t max_region_gap, int32_t max_cluster_gap, int basic,
                         dfb_regions** out);
void dfb_regions_free(dfb_regions* r);
int64_t dfb_regions_count(const dfb_regions* r);
int64_t dfb_regions_count_up(const dfb_regions* r); /* rows named "up"; the rest are "dn" */
/* Any output pointer may be NULL.  start/end genomic (full mode only), value = mean, area =
 * |sum|, cluster ids, clusterL (NaN encodes R's NA).  Rows: all "up" then all "dn" (:292). */
int dfb_regions_get(const dfb_regions* r, int32_t* start, int32_t* end, double* value,
                    double* area, int32_t* index_start, int32_t* index_end, int32_t* cluster,
                    double* cluster_l);
/* mean(Views(v, IRanges(indexStart, indexEnd))) for a host vector v[N] -- R/calculatePvalues.R:
 * 253-262 (meanCoverage / group means per region). which: -1 = cov meanCoverage, g>=0 = group g */
int dfb_regions_view_means(dfb_ctx* ctx, const dfb_regions* r, const dfb_cov* cov, int which,
                           double* out);
/* The same for a caller-supplied host vector v[N] (a coveragePrep$meanCoverage / groupMeans entry
 * that did not come from this handle; the reference uses whatever vector it is given,
 * R/calculatePvalues.R:222-224, 253-262). */
int dfb_regions_view_means_host(dfb_ctx* ctx, const dfb_regions* r, const double* v, int64_t N,
                                double* out);

/* .clusterMakerRle() -- R/findRegions.R:445-475 on a logical Rle given as run lengths.
 * ranges = 0: fills ids[k] = k+1 and counts[k] (the integer Rle); ranges = 1: fills
 * starts/ends (index space).  Returns the number of clusters; pass NULL outputs to size. */
int64_t dfb_cluster_maker(const int32_t* pos_len, int64_t pos_nruns, int pos_first, int32_t max_gap,
                          int64_t* counts, int64_t* starts, int64_t* ends, int64_t cap);

/* ---------------------------------------------------------------- permutation nulls --------- */

/* The permutation loop of calculatePvalues() -- R/calculatePvalues.R:306-358.
 * perm_idx [B x n] row-major, 0-based: row b is `sample(nSamples) - 1` drawn by the R glue under
 * `set.seed(seeds[b])` (:315-318).  For every permutation the engine computes the F-stats of
 * the row-permuted models, thresholds them, finds the candidate regions (basic=TRUE) inside the
 * position segments and keeps {stat, width, area, permutation}.  Null F-stats never touch HBM
 * beyond a bit mask and the exceedance values. */
int dfb_perm_nulls(dfb_ctx* ctx, dfb_cov* cov, const double* mod, int p, const double* mod0,
                   int p0, double adjust_f, const int32_t* perm_idx, int64_t B, double cutoff_lo,
                   double cutoff_hi, int32_t max_region_gap, dfb_nulls** out);
void dfb_nulls_free(dfb_nulls* nl);
/* number of DISTINCT null regions (the reference stores each twice, :346-354; see dup below) */
int64_t dfb_nulls_count(const dfb_nulls* nl);
/* per-permutation region counts [B] */
int dfb_nulls_counts_per_perm(const dfb_nulls* nl, int64_t* counts);
/* dup = 0: each region once, ordered by (permutation, up/dn, index).
 * dup = 1: the reference's layout, 2*count rows: per permutation [r1..rk, r1..rk]. */
int dfb_nulls_get(const dfb_nulls* nl, int dup, double* stat, int32_t* width, double* area,
                  int32_t* permutation);
/* device access for multi-GPU pooling (areas as doubles, permutation ids as int32) */
int dfb_nulls_copy_dev(const dfb_nulls* nl, double* area_dev, int32_t* perm_dev);
/* Genome-wide ranking of mergeResults() across GPUs (R/mergeResults.R:216-235, 300-307, 380-386) in
 * three calls per rank around two small collectives.  Ranks never exchange their null lists: they
 * all-gather their REGION areas, histogram their own nulls against the merged region list and
 * all-reduce that histogram, so per-rank work and traffic do not grow with the number of GPUs.
 *   dfb_pool_local_dev  record [2 + B + rcap] = {R, M, max null area of permutation 1..B (-1 when
 *                       the rank has none), this rank's region areas, descending, padded with -1};
 *                       merge_area != 0: null areas are abs(stat) * width (:231).  A rank with
 *                       R > rcap writes the first rcap areas; every rank then sees R > rcap in
 *                       the gathered records and the exchange is repeated with a larger capacity.
 *                                                                 -> all-gather the records
 *   dfb_pool_hist_dev   merged_dev / src_dev [world * rcap]: every rank's region areas merged into
 *                       one descending list and where each came from (rank * rcap + position);
 *                       hist_dev [world * rcap + 1]: this rank's nulls binned by the number of
 *                       pooled regions they do not exceed.        -> all-reduce(sum) hist_dev
 *   dfb_pool_rank_dev   pooled p-values ((2 * #{null > area} + 1) / (2 * sum(M) + 1), i.e. .calcPval
 *                       on the duplicated pool) and FWER (.calculateFWER, empty permutation ->
 *                       cutoff_used) of THIS rank's regions, [R] on the device in up-then-dn order.
 * A chromosome without observed regions (calculatePvalues returned NULLs, :245-251) passes r = nl = NULL
 * to all three calls: it contributes an empty record and no nulls. */
int dfb_pool_local_dev(dfb_ctx* ctx, const dfb_regions* r, dfb_nulls* nl, int64_t B, int merge_area,
                       int64_t rcap, double* record_dev);
int dfb_pool_hist_dev(dfb_ctx* ctx, dfb_nulls* nl, int merge_area,
                      const double* records_dev /* [world][2 + B + rcap] */, int world, int64_t B,
                      int64_t rcap, double* merged_dev, int32_t* src_dev, unsigned long long* hist_dev);
int dfb_pool_rank_dev(dfb_ctx* ctx, const dfb_regions* r, const double* records_dev, int world,
                      int rank, int64_t B, int64_t rcap, const double* merged_dev,
                      const int32_t* src_dev, const unsigned long long* hist_dev, double cutoff_used,
                      double* p_dev, double* fwer_dev);
/* the null areas as mergeResults() pools them: abs(stat) * width (R/mergeResults.R:231), which is
 * not bit-identical to the summed area calculatePvalues ranks; written to device memory [count] */
int dfb_nulls_merge_areas_dev(const dfb_nulls* nl, double* area_dev);

/* BigWig data sections -> coverage of the regions, on the device (R/railMatrix.R:271-289: per sample
 * `import(sampleFile, selection = regs, as = "RleList")`, `/ mappedPerXM`; the view sums follow with
 * dfb_view_sums).  The host finds the file's leaf blocks that overlap the regions (R-tree), inflates them
 * and passes, per sample s, the concatenated uncompressed data sections `sections[s]` with their byte
 * offsets `sec_off[s][0 .. nsec[s]]` and the id the file gives the chromosome (`chrom_id[s]`).  Regions:
 * 1-based inclusive, sorted, disjoint.  Result: an F64 coverage handle of N = sum(width(regions)) bases
 * (region k occupies rows sum(width[0..k)) + 1 ...), value / mapped_per_xm[s] where the file has data,
 * 0 elsewhere. */
int dfb_cov_from_bigwig(dfb_ctx* ctx, int n, const void* const* sections, const int64_t* const* sec_off,
                        const int64_t* nsec, const uint32_t* chrom_id, const int32_t* rstart,
                        const int32_t* rend, int64_t R, const double* mapped_per_xm, dfb_cov** out);

/* ---------------------------------------------------------------- mergeResults -------------- */

/* One-call genome-wide pass of mergeResults() -- R/mergeResults.R:216-235 (pool the null regions of every
 * chromosome, area = abs(stat) * width when merge_area != 0, :231), :300-307 (.calcPval of every region
 * against the pool) and :380-386 (.calculateFWER: per-permutation maximum over the genome, an empty
 * permutation counts as cutoff_used) -- across all GPUs of the job.  Every rank passes the results of the
 * chromosomes IT holds (nchr >= 0; a chromosome without observed regions passes r[c] = nl[c] = NULL, see
 * R/calculatePvalues.R:245-251); the ranks exchange their region areas and per-permutation maxima
 * (all-gather) and the histogram of their nulls over the merged region list (all-reduce) on the context's
 * communicator -- the null lists never leave their GPU.  Outputs, per chromosome c, in the order of
 * dfb_regions_get (decreasing area): pvalues[c][R_c], fwer[c][R_c] (host; either array of pointers may be
 * NULL); *total_nulls = number of distinct null regions in the genome-wide pool.  Collective: every rank of
 * the communicator must call it (with the same B, cutoff_used and merge_area).  A context without a
 * communicator pools its own chromosomes only. */
#define DFB_COMM_ID_BYTES 128
int dfb_comm_unique_id(void* id_out /* DFB_COMM_ID_BYTES */);
int dfb_comm_init_rank(dfb_ctx* ctx, int world, int rank, const void* id /* DFB_COMM_ID_BYTES */);
int dfb_comm_init_all(dfb_ctx* const* ctxs, int ngpu); /* one process, one context per GPU */
void dfb_comm_destroy(dfb_ctx* ctx);
int dfb_comm_info(const dfb_ctx* ctx, int* world, int* rank);
int dfb_merge_results(dfb_ctx* ctx, int nchr, dfb_regions* const* r, dfb_nulls* const* nl, int64_t B,
                      double cutoff_used, int merge_area, double* const* pvalues, double* const* fwer,
                      int64_t* total_nulls);

/* Multi-GPU pooling without a second sort (R/mergeResults.R:216-235, 304-307): every rank keeps its
 * null areas sorted from its own p-value ranking; ranks all-gather those sorted lists and count,
 * list by list, how many pooled nulls exceed each of their region areas:
 *   p = (2 * sum_ranks #{null > area} + 1) / (2 * sum_ranks M + 1).
 * dfb_nulls_sorted_areas_dev copies the ascending null areas (count doubles) to out_dev;
 * dfb_count_greater_dev ADDS #{sorted_nulls > areas[r]} to counts_dev[r];
 * dfb_regions_copy_areas_dev copies the region areas in up-then-dn order (before the
 * decreasing-area reordering, see dfb_regions_get_order). */
int dfb_nulls_sorted_areas_dev(dfb_ctx* ctx, dfb_nulls* nl, double* out_dev);
/* Ranking without sorting any nulls (what dfb_calculate_pvalues does itself): the regions of `r`
 * are already sorted by decreasing area, so a null v exceeds exactly the regions from position
 * #{area >= v} on.  dfb_regions_null_hist_dev ADDS the histogram of those positions for M unsorted
 * nulls to hist_dev [R + 1] (uint32, zeroed by the caller; call once per rank's list);
 * dfb_regions_hist_counts_dev turns it into counts_dev[k] = #{pooled null > area of region k},
 * regions in up-then-dn order. */
int dfb_regions_null_hist_dev(dfb_ctx* ctx, const dfb_regions* r, const double* nulls_dev, int64_t M,
                              uint32_t* hist_dev);
int dfb_regions_hist_counts_dev(dfb_ctx* ctx, const dfb_regions* r, const uint32_t* hist_dev,
                                int64_t* counts_dev);
int dfb_count_greater_dev(dfb_ctx* ctx, const double* areas_dev, int64_t R, const double* sorted_nulls_dev,
                          int64_t M, int64_t* counts_dev);
int dfb_regions_copy_areas_dev(dfb_ctx* ctx, const dfb_regions* r, double* out_dev);

/* Validation hook for the tcgen05 screening path (used by the parity tests and profiling): runs
 * the dense FP64 scan and the screen + refinement on the same permutations with a scalar cutoff
 * and reports the exceedance counts of both, the number of screened candidates and the number of
 * mask bits that differ between the dense path and the screened paths (two-kernel and, when
 * supported, fused; must be 0). */
int dfb_debug_screen(dfb_ctx* ctx, dfb_cov* cov, const double* mod, int p, const double* mod0, int p0,
                     double adjust_f, const int32_t* perm_idx, int64_t B, double cutoff,
                     int64_t* n_exact_dense, int64_t* n_exact_screen, int64_t* n_candidates,
                     int64_t* n_mask_diff);

/* ---------------------------------------------------------------- p-values ------------------ */

/* .calcPval() -- R/calculatePvalues.R:432-441.  weight = multiplicity of every null entry
 * (2 reproduces the duplicated null list): p = (weight*#{null > a*} + 1) / (weight*M + 1) with
 * a* the largest null <= area (all nulls when none is). */
int dfb_calc_pval(dfb_ctx* ctx, const double* areas, int64_t R, const double* nullareas,
                  int64_t M, int weight, double* pvalues);
int dfb_calc_pval_dev(dfb_ctx* ctx, const double* areas_dev, int64_t R, double* nullareas_dev,
                      int64_t M, int weight, double* pvalues_dev);
/* .calculateFWER() -- R/mergeResults.R:380-386. permutation ids 1..B. */
int dfb_calc_fwer(dfb_ctx* ctx, double cutoff_used, const double* nullareas,
                  const int32_t* permutation, int64_t M, int64_t B, const double* areas, int64_t R,
                  double* fwer);
/* per-permutation maxima [B] on the device (input to an all-reduce(max) across GPUs) */
int dfb_perm_max_dev(dfb_ctx* ctx, const double* nullareas_dev, const int32_t* perm_dev, int64_t M,
                     int64_t B, double fill, double* max_dev);
/* quantile(x, prob, na.rm=TRUE), R type 7 -- default cutoff of findRegions/calculatePvalues
 * (R/findRegions.R:118, R/analyzeChr.R:318).  x = NULL uses the cached F-stats of cov. */
int dfb_quantile(dfb_ctx* ctx, const dfb_cov* cov, const double* x, int64_t N, double prob,
                 double* out);

/* calculatePvalues() end to end -- R/calculatePvalues.R:157-430 minus qvalue():
 * observed regions (full mode) sorted by decreasing area (stable, :368), their meanCoverage /
 * group means, null regions, p-values.  Returns DFB_ENOREGIONS when no observed region. */
int dfb_calculate_pvalues(dfb_ctx* ctx, dfb_cov* cov, const double* fstats, const double* mod,
                          int p, const double* mod0, int p0, double adjust_f,
                          const int32_t* perm_idx, int64_t B, double cutoff_lo, double cutoff_hi,
                          int32_t max_region_gap, int32_t max_cluster_gap, dfb_regions** regions,
                          dfb_nulls** nulls, double* pvalues /* [R], may be NULL: fetch later */);
int dfb_regions_get_pvalues(const dfb_regions* r, double* pvalues);
/* row order applied by dfb_calculate_pvalues (0-based indices into the up-then-dn order) */
int dfb_regions_get_order(const dfb_regions* r, int32_t* order);

/* ---------------------------------------------------------------- ER path ------------------- */

/* coverageMatrix of regionMatrix() -- R/regionMatrix.R:252-270 with getRegionCoverage
 * (R/getRegionCoverage.R:135-160): out[r + R*s] = sum of rows indexStart..indexEnd of column s.
 * The division by L (scalar, or the reference's last-column-only vector quirk :267) is applied
 * when L != NULL (n_l = 1 or n). Column-major [R x n] like an R matrix. */
int dfb_region_matrix(dfb_ctx* ctx, const dfb_cov* cov, const int32_t* index_start,
                      const int32_t* index_end, int64_t R, const double* L, int n_l, double* out);
/* same, regions taken from a dfb_regions result (stays on the device until the final copy) */
int dfb_region_matrix_r(dfb_ctx* ctx, const dfb_cov* cov, const dfb_regions* r, const double* L,
                        int n_l, double* out);

/* railMatrix()'s per-sample column -- .railMatChrRegionCov (R/railMatrix.R:264-290):
 * out[r + R*s] = sum(Views(cov_s, IRanges(start, end)))[r] / L_s over rows start..end (1-based,
 * inclusive) of column s.  Library-size normalised (double) coverage is summed in IRanges' Rle
 * order -- one product value * length per run of identical values, left to right -- so the result
 * is bit-identical to `regCov / mappedPerXM` followed by viewSums; integer coverage is exact.
 * L: NULL, one value, or one per sample (every column is divided, unlike regionMatrix's quirk). */
int dfb_view_sums(dfb_ctx* ctx, const dfb_cov* cov, const int32_t* start, const int32_t* end,
                  int64_t R, const double* L, int n_l, double* out);

#ifdef __cplusplus
}
#endif
#endif /* DERFINDER_B200_H */
