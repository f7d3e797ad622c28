/*
 * .Call shim between R and libderfinder_b200.so  (see INTEGRATION.md).
 *
 * Pure marshalling: every function unpacks SEXPs into the plain pointers / sizes of
 * include/derfinder_b200.h, calls ONE C-ABI entry point (or a fetch sequence), and wraps results in
 * freshly allocated R vectors.  Device objects are R external pointers with finalisers, so an
 * Rf_error() longjmp never leaks device memory: nothing device-side is owned by the C stack.
 *
 * NOT compiled in this repository's CI image (no R headers there); build inside an R package with
 *   PKG_CPPFLAGS = -I<repo>/include      PKG_LIBS = -L<repo>/derfinder_b200 -lderfinder_b200
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <stdint.h>
#include <string.h>

#include "derfinder_b200.h"

#define CHECK(expr)                                     \
    do {                                                \
        int _rc = (expr);                               \
        if (_rc < 0) Rf_error("%s", dfb_last_error());  \
    } while (0)

/* ---------------------------------------------------------------- handles ------------------- */

static void ctx_finalizer(SEXP p) {
    dfb_ctx* c = (dfb_ctx*)R_ExternalPtrAddr(p);
    if (c) dfb_destroy(c);
    R_ClearExternalPtr(p);
}
static void cov_finalizer(SEXP p) {
    dfb_cov* c = (dfb_cov*)R_ExternalPtrAddr(p);
    if (c) dfb_cov_free(c);
    R_ClearExternalPtr(p);
}
static void regions_finalizer(SEXP p) {
    dfb_regions* r = (dfb_regions*)R_ExternalPtrAddr(p);
    if (r) dfb_regions_free(r);
    R_ClearExternalPtr(p);
}
static void nulls_finalizer(SEXP p) {
    dfb_nulls* n = (dfb_nulls*)R_ExternalPtrAddr(p);
    if (n) dfb_nulls_free(n);
    R_ClearExternalPtr(p);
}
/* Child handles (coverage, regions, nulls) keep their context alive: the context's external pointer is the
 * `prot` value of theirs, so the garbage collector cannot finalise it first, and no finaliser runs at exit
 * (onexit = FALSE) where R gives no order at all -- the process teardown releases the device.  dfb_cov_free /
 * dfb_nulls_free / dfb_regions_free dereference ctx->stream, so the context must outlive them. */
static SEXP wrap(void* h, const char* tag, R_CFinalizer_t fin, SEXP owner) {
    SEXP p = PROTECT(R_MakeExternalPtr(h, Rf_install(tag), owner));
    R_RegisterCFinalizerEx(p, fin, FALSE);
    UNPROTECT(1);
    return p;
}
static void* unwrap(SEXP p, const char* what) {
    void* h = R_ExternalPtrAddr(p);
    if (!h) Rf_error("%s: stale device handle (object was saved/loaded or freed)", what);
    return h;
}

SEXP dfbR_create(SEXP device) {
    dfb_ctx* c = NULL;
    CHECK(dfb_create(Rf_asInteger(device), NULL, &c));
    return wrap(c, "dfb_ctx", ctx_finalizer, R_NilValue);
}

/* ---------------------------------------------------------------- Rle marshalling ----------- */
/* A DataFrame of Rle columns arrives as two lists: values (integer or numeric vectors) and
 * lengths (integer vectors), i.e. lapply(df, runValue) / lapply(df, runLength). */
static int rle_columns(SEXP values, SEXP lengths, const void*** v, const int32_t*** l, int64_t** nr) {
    int n = Rf_length(values);
    if (n < 1 || Rf_length(lengths) != n) Rf_error("coverage must have >= 1 Rle column");
    int dtype = TYPEOF(VECTOR_ELT(values, 0)) == INTSXP ? DFB_I32 : DFB_F64;
    *v = (const void**)R_alloc(n, sizeof(void*));
    *l = (const int32_t**)R_alloc(n, sizeof(int32_t*));
    *nr = (int64_t*)R_alloc(n, sizeof(int64_t));
    for (int s = 0; s < n; ++s) {
        SEXP vs = VECTOR_ELT(values, s), ls = VECTOR_ELT(lengths, s);
        if ((TYPEOF(vs) == INTSXP ? DFB_I32 : DFB_F64) != dtype) Rf_error("mixed integer/numeric coverage columns");
        (*v)[s] = dtype == DFB_I32 ? (const void*)INTEGER(vs) : (const void*)REAL(vs);
        (*l)[s] = (const int32_t*)INTEGER(ls);
        (*nr)[s] = XLENGTH(vs);
    }
    return dtype;
}

/* filterData(): values/lengths lists, chromosome length, mappedPerXM (or NULL), filter kind,
 * cutoff, previous index as run lengths (or NULL) + value of its first run. */
SEXP dfbR_filter(SEXP ctx, SEXP values, SEXP lengths, SEXP len, SEXP mapped, SEXP filter, SEXP cutoff,
                 SEXP prev_len, SEXP prev_first) {
    const void** v; const int32_t** l; int64_t* nr;
    int dtype = rle_columns(values, lengths, &v, &l, &nr);
    dfb_cov* cov = NULL;
    CHECK(dfb_filter_rle((dfb_ctx*)unwrap(ctx, "filterData"), Rf_length(values), dtype, v, l, nr,
                         (int64_t)Rf_asReal(len), Rf_isNull(mapped) ? NULL : REAL(mapped), Rf_asInteger(filter),
                         Rf_asReal(cutoff), Rf_isNull(prev_len) ? NULL : (const int32_t*)INTEGER(prev_len),
                         Rf_isNull(prev_len) ? 0 : XLENGTH(prev_len), Rf_asLogical(prev_first), &cov));
    return wrap(cov, "dfb_cov", cov_finalizer, ctx);
}

/* already filtered DataFrame (what loadCoverage(cutoff=) returns) + its position index */
SEXP dfbR_cov_from_rle(SEXP ctx, SEXP values, SEXP lengths, SEXP N, SEXP pos_len, SEXP pos_first) {
    const void** v; const int32_t** l; int64_t* nr;
    int dtype = rle_columns(values, lengths, &v, &l, &nr);
    dfb_cov* cov = NULL;
    CHECK(dfb_cov_from_rle((dfb_ctx*)unwrap(ctx, "coverage"), Rf_length(values), dtype, v, l, nr,
                           (int64_t)Rf_asReal(N), (const int32_t*)INTEGER(pos_len), XLENGTH(pos_len),
                           Rf_asLogical(pos_first), &cov));
    return wrap(cov, "dfb_cov", cov_finalizer, ctx);
}

SEXP dfbR_cov_info(SEXP cov) {
    int64_t N, chr_len, pos_nruns; int n, dtype, G, has;
    CHECK(dfb_cov_info((dfb_cov*)unwrap(cov, "coverage"), &N, &n, &chr_len, &dtype, &pos_nruns, &G, &has));
    SEXP out = PROTECT(Rf_allocVector(REALSXP, 7));
    double* o = REAL(out);
    o[0] = (double)N; o[1] = n; o[2] = (double)chr_len; o[3] = dtype; o[4] = (double)pos_nruns; o[5] = G; o[6] = has;
    UNPROTECT(1);
    return out;
}

/* position as list(lengths, first) -> the glue builds Rle(rep_len(c(first, !first), k), lengths) */
SEXP dfbR_cov_position(SEXP cov) {
    dfb_cov* c = (dfb_cov*)unwrap(cov, "coverage");
    int64_t N, chr_len, k; int n, dtype, G, has, first = 0;
    CHECK(dfb_cov_info(c, &N, &n, &chr_len, &dtype, &k, &G, &has));
    SEXP len = PROTECT(Rf_allocVector(INTSXP, (R_xlen_t)k));
    CHECK(dfb_cov_get_position(c, (int32_t*)INTEGER(len), &first));
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));
    SET_VECTOR_ELT(out, 0, len);
    SET_VECTOR_ELT(out, 1, Rf_ScalarLogical(first));
    UNPROTECT(2);
    return out;
}

/* which: 0 raw column, 1 log2 column, 2 meanCoverage, 3 group mean; idx = sample / group (0-based) */
SEXP dfbR_cov_vector(SEXP cov, SEXP which, SEXP idx) {
    dfb_cov* c = (dfb_cov*)unwrap(cov, "coverage");
    int64_t N, chr_len, k; int n, dtype, G, has;
    CHECK(dfb_cov_info(c, &N, &n, &chr_len, &dtype, &k, &G, &has));
    int w = Rf_asInteger(which), i = Rf_asInteger(idx);
    SEXP out;
    if (w == 0 && dtype == DFB_I32) {
        out = PROTECT(Rf_allocVector(INTSXP, (R_xlen_t)N));
        CHECK(dfb_cov_get_column(c, i, INTEGER(out)));
    } else {
        out = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)N));
        if (w == 0) CHECK(dfb_cov_get_column(c, i, REAL(out)));
        else if (w == 1) CHECK(dfb_cov_get_processed_column(c, i, REAL(out)));
        else if (w == 2) CHECK(dfb_cov_get_mean(c, REAL(out)));
        else CHECK(dfb_cov_get_group_mean(c, i, REAL(out)));
    }
    UNPROTECT(1);
    return out;
}

SEXP dfbR_preprocess(SEXP ctx, SEXP cov, SEXP scalefac, SEXP group_of_sample, SEXP n_groups) {
    CHECK(dfb_preprocess((dfb_ctx*)unwrap(ctx, "preprocessCoverage"), (dfb_cov*)unwrap(cov, "coverage"),
                         Rf_asReal(scalefac),
                         Rf_isNull(group_of_sample) ? NULL : (const int32_t*)INTEGER(group_of_sample),
                         Rf_asInteger(n_groups)));
    return R_NilValue;
}

SEXP dfbR_chunk_lengths(SEXP N, SEXP chunksize, SEXP cores) {
    int64_t k = dfb_chunk_lengths((int64_t)Rf_asReal(N), Rf_asReal(chunksize), Rf_asInteger(cores), NULL, 0);
    int64_t* tmp = (int64_t*)R_alloc((size_t)(k > 0 ? k : 1), sizeof(int64_t));
    dfb_chunk_lengths((int64_t)Rf_asReal(N), Rf_asReal(chunksize), Rf_asInteger(cores), tmp, k);
    SEXP out = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)k));
    for (int64_t i = 0; i < k; ++i) REAL(out)[i] = (double)tmp[i];
    UNPROTECT(1);
    return out;
}

/* ---------------------------------------------------------------- F statistics -------------- */

SEXP dfbR_fstats(SEXP ctx, SEXP cov, SEXP mod, SEXP mod0, SEXP adjustF) {
    dfb_cov* c = (dfb_cov*)unwrap(cov, "calculateStats");
    int64_t N, chr_len, k; int n, dtype, G, has;
    CHECK(dfb_cov_info(c, &N, &n, &chr_len, &dtype, &k, &G, &has));
    if (Rf_nrows(mod) != n || Rf_nrows(mod0) != n)
        Rf_error("The alternative model matrix is not conformable with the data matrix.");
    SEXP out = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)N));
    CHECK(dfb_fstats((dfb_ctx*)unwrap(ctx, "calculateStats"), c, REAL(mod), Rf_ncols(mod), REAL(mod0), Rf_ncols(mod0),
                     Rf_asReal(adjustF), REAL(out)));
    UNPROTECT(1);
    return out;
}

/* ---------------------------------------------------------------- regions ------------------- */

static SEXP regions_to_list(dfb_regions* r, int basic) {
    R_xlen_t R = (R_xlen_t)dfb_regions_count(r);
    const char* nm[] = {"start", "end", "value", "area", "indexStart", "indexEnd", "cluster", "clusterL", "nUp", ""};
    SEXP out = PROTECT(Rf_mkNamed(VECSXP, nm));
    SEXP st = PROTECT(Rf_allocVector(INTSXP, basic ? 0 : R)), en = PROTECT(Rf_allocVector(INTSXP, basic ? 0 : R));
    SEXP val = PROTECT(Rf_allocVector(REALSXP, R)), area = PROTECT(Rf_allocVector(REALSXP, R));
    SEXP is = PROTECT(Rf_allocVector(INTSXP, R)), ie = PROTECT(Rf_allocVector(INTSXP, R));
    SEXP cl = PROTECT(Rf_allocVector(INTSXP, basic ? 0 : R)), cll = PROTECT(Rf_allocVector(REALSXP, basic ? 0 : R));
    CHECK(dfb_regions_get(r, basic ? NULL : (int32_t*)INTEGER(st), basic ? NULL : (int32_t*)INTEGER(en), REAL(val),
                          REAL(area), (int32_t*)INTEGER(is), (int32_t*)INTEGER(ie),
                          basic ? NULL : (int32_t*)INTEGER(cl), basic ? NULL : REAL(cll)));
    if (!basic)
        for (R_xlen_t i = 0; i < R; ++i)
            if (ISNAN(REAL(cll)[i])) REAL(cll)[i] = NA_REAL; /* NaN encodes R's NA (:266-271) */
    SET_VECTOR_ELT(out, 0, st); SET_VECTOR_ELT(out, 1, en); SET_VECTOR_ELT(out, 2, val);
    SET_VECTOR_ELT(out, 3, area); SET_VECTOR_ELT(out, 4, is); SET_VECTOR_ELT(out, 5, ie);
    SET_VECTOR_ELT(out, 6, cl); SET_VECTOR_ELT(out, 7, cll);
    SET_VECTOR_ELT(out, 8, Rf_ScalarReal((double)dfb_regions_count_up(r)));
    UNPROTECT(9);
    return out;
}

/* findRegions(): x = fstats (numeric) ; position as run lengths + first value */
SEXP dfbR_find_regions(SEXP ctx, SEXP x, SEXP pos_len, SEXP pos_first, SEXP cut_lo, SEXP cut_hi, SEXP max_region_gap,
                       SEXP max_cluster_gap, SEXP basic) {
    dfb_regions* r = NULL;
    int rc = dfb_find_regions_pos((dfb_ctx*)unwrap(ctx, "findRegions"), REAL(x), XLENGTH(x),
                                  (const int32_t*)INTEGER(pos_len), XLENGTH(pos_len), Rf_asLogical(pos_first),
                                  Rf_asReal(cut_lo), Rf_asReal(cut_hi), Rf_asInteger(max_region_gap),
                                  Rf_asInteger(max_cluster_gap), Rf_asLogical(basic), &r);
    CHECK(rc);
    if (rc == DFB_ENOREGIONS) return R_NilValue; /* glue issues warning("Found no regions") and returns NULL */
    SEXP h = PROTECT(wrap(r, "dfb_regions", regions_finalizer, ctx)); /* freed by the GC, also on error */
    (void)h;
    SEXP out = regions_to_list(r, Rf_asLogical(basic));
    UNPROTECT(1);
    return out;
}

/* ---------------------------------------------------------------- calculatePvalues ---------- */

/* fstats = NULL uses the statistics cached on the device by dfbR_fstats (the glue passes NULL when
 * the `fstats` argument still carries the handle attribute calculateStats() attached).
 * perm: integer matrix [n x B] = sapply(seeds, function(s) { set.seed(s); sample(n) }) - 1L. */
/* mean_v / group_v: NULL = summarise the device copies (the vectors preprocessCoverage() returned for this handle);
 * otherwise the caller's own meanCoverage (numeric) / groupMeans (list of numeric), R/calculatePvalues.R:222-262 */
SEXP dfbR_calculate_pvalues(SEXP ctx, SEXP cov, SEXP fstats, SEXP mod, SEXP mod0, SEXP adjustF, SEXP perm, SEXP cut_lo,
                            SEXP cut_hi, SEXP max_region_gap, SEXP max_cluster_gap, SEXP n_groups, SEXP mean_v,
                            SEXP group_v) {
    dfb_ctx* cx = (dfb_ctx*)unwrap(ctx, "calculatePvalues");
    dfb_cov* c = (dfb_cov*)unwrap(cov, "calculatePvalues");
    dfb_regions* r = NULL;
    dfb_nulls* nl = NULL;
    /* R matrices are column-major: an [n x B] integer matrix IS perm_idx[B][n] row-major */
    int rc = dfb_calculate_pvalues(cx, c, Rf_isNull(fstats) ? NULL : REAL(fstats), REAL(mod), Rf_ncols(mod), REAL(mod0),
                                   Rf_ncols(mod0), Rf_asReal(adjustF), (const int32_t*)INTEGER(perm), Rf_ncols(perm),
                                   Rf_asReal(cut_lo), Rf_asReal(cut_hi), Rf_asInteger(max_region_gap),
                                   Rf_asInteger(max_cluster_gap), &r, &nl, NULL);
    CHECK(rc);
    if (rc == DFB_ENOREGIONS) return R_NilValue; /* list(regions = NULL, nullStats = NULL, ...) (:245-251) */
    SEXP hr = PROTECT(wrap(r, "dfb_regions", regions_finalizer, ctx));
    SEXP hn = PROTECT(wrap(nl, "dfb_nulls", nulls_finalizer, ctx));
    /* owned by the GC from here on: an Rf_error below cannot leak them */
    const char* nm[] = {"regions", "pvalues", "meanCoverage", "groupMeans", "nullStats", "nullWidths",
                        "nullPermutation", "nullAreas", "handles", ""};
    SEXP out = PROTECT(Rf_mkNamed(VECSXP, nm));
    {
        const char* hn_names[] = {"regions", "nulls", ""};
        SEXP hl = PROTECT(Rf_mkNamed(VECSXP, hn_names));
        SET_VECTOR_ELT(hl, 0, hr);
        SET_VECTOR_ELT(hl, 1, hn);
        SET_VECTOR_ELT(out, 8, hl);
        UNPROTECT(1);
    }
    SET_VECTOR_ELT(out, 0, regions_to_list(r, 0));
    R_xlen_t R = (R_xlen_t)dfb_regions_count(r), M = (R_xlen_t)dfb_nulls_count(nl);
    SEXP pv = PROTECT(Rf_allocVector(REALSXP, R));
    CHECK(dfb_regions_get_pvalues(r, REAL(pv)));
    if (M == 0) for (R_xlen_t i = 0; i < R; ++i) REAL(pv)[i] = NA_REAL; /* :408-419 */
    SET_VECTOR_ELT(out, 1, pv);
    SEXP mc = PROTECT(Rf_allocVector(REALSXP, R));
    if (Rf_isNull(mean_v)) CHECK(dfb_regions_view_means(cx, r, c, -1, REAL(mc)));
    else CHECK(dfb_regions_view_means_host(cx, r, REAL(mean_v), XLENGTH(mean_v), REAL(mc)));
    SET_VECTOR_ELT(out, 2, mc);
    int G = Rf_asInteger(n_groups);
    SEXP gm = PROTECT(Rf_allocVector(VECSXP, G));
    for (int g = 0; g < G; ++g) {
        SEXP v = PROTECT(Rf_allocVector(REALSXP, R));
        if (Rf_isNull(group_v)) CHECK(dfb_regions_view_means(cx, r, c, g, REAL(v)));
        else CHECK(dfb_regions_view_means_host(cx, r, REAL(VECTOR_ELT(group_v, g)), XLENGTH(VECTOR_ELT(group_v, g)), REAL(v)));
        SET_VECTOR_ELT(gm, g, v);
        UNPROTECT(1);
    }
    SET_VECTOR_ELT(out, 3, gm);
    SEXP ns = PROTECT(Rf_allocVector(REALSXP, 2 * M)), nw = PROTECT(Rf_allocVector(INTSXP, 2 * M));
    SEXP np = PROTECT(Rf_allocVector(INTSXP, 2 * M)), na = PROTECT(Rf_allocVector(REALSXP, 2 * M));
    CHECK(dfb_nulls_get(nl, 1, REAL(ns), (int32_t*)INTEGER(nw), REAL(na), (int32_t*)INTEGER(np)));
    SET_VECTOR_ELT(out, 4, ns); SET_VECTOR_ELT(out, 5, nw); SET_VECTOR_ELT(out, 6, np); SET_VECTOR_ELT(out, 7, na);
    UNPROTECT(10);
    return out;
}

SEXP dfbR_calc_pval(SEXP ctx, SEXP areas, SEXP nullareas) {
    SEXP out = PROTECT(Rf_allocVector(REALSXP, XLENGTH(areas)));
    CHECK(dfb_calc_pval((dfb_ctx*)unwrap(ctx, ".calcPval"), REAL(areas), XLENGTH(areas), REAL(nullareas),
                        XLENGTH(nullareas), 1, REAL(out)));
    UNPROTECT(1);
    return out;
}

SEXP dfbR_calc_fwer(SEXP ctx, SEXP cutoff_used, SEXP nullareas, SEXP permutation, SEXP nperm, SEXP areas) {
    SEXP out = PROTECT(Rf_allocVector(REALSXP, XLENGTH(areas)));
    CHECK(dfb_calc_fwer((dfb_ctx*)unwrap(ctx, ".calculateFWER"), Rf_asReal(cutoff_used), REAL(nullareas),
                        (const int32_t*)INTEGER(permutation), XLENGTH(nullareas), (int64_t)Rf_asInteger(nperm),
                        REAL(areas), XLENGTH(areas), REAL(out)));
    UNPROTECT(1);
    return out;
}

SEXP dfbR_quantile(SEXP ctx, SEXP x, SEXP prob) {
    double q = NA_REAL;
    CHECK(dfb_quantile((dfb_ctx*)unwrap(ctx, "quantile"), NULL, REAL(x), XLENGTH(x), Rf_asReal(prob), &q));
    return Rf_ScalarReal(q);
}

/* ---------------------------------------------------------------- regionMatrix -------------- */

SEXP dfbR_region_matrix(SEXP ctx, SEXP cov, SEXP index_start, SEXP index_end, SEXP L) {
    dfb_cov* c = (dfb_cov*)unwrap(cov, "regionMatrix");
    int64_t N, chr_len, k; int n, dtype, G, has;
    CHECK(dfb_cov_info(c, &N, &n, &chr_len, &dtype, &k, &G, &has));
    R_xlen_t R = XLENGTH(index_start);
    SEXP out = PROTECT(Rf_allocMatrix(REALSXP, (int)R, n));
    CHECK(dfb_region_matrix((dfb_ctx*)unwrap(ctx, "regionMatrix"), c, (const int32_t*)INTEGER(index_start),
                            (const int32_t*)INTEGER(index_end), R, Rf_isNull(L) ? NULL : REAL(L),
                            Rf_isNull(L) ? 0 : Rf_length(L), REAL(out)));
    UNPROTECT(1);
    return out;
}

/* railMatrix: per-sample sum(Views(cov, ranges)) / L  (R/railMatrix.R:264-290) */
SEXP dfbR_view_sums(SEXP ctx, SEXP cov, SEXP start, SEXP end, SEXP L) {
    dfb_cov* c = (dfb_cov*)unwrap(cov, "railMatrix");
    int64_t N, chr_len, k; int n, dtype, G, has;
    CHECK(dfb_cov_info(c, &N, &n, &chr_len, &dtype, &k, &G, &has));
    R_xlen_t R = XLENGTH(start);
    SEXP out = PROTECT(Rf_allocMatrix(REALSXP, (int)R, n));
    CHECK(dfb_view_sums((dfb_ctx*)unwrap(ctx, "railMatrix"), c, (const int32_t*)INTEGER(start),
                        (const int32_t*)INTEGER(end), R, Rf_isNull(L) ? NULL : REAL(L),
                        Rf_isNull(L) ? 0 : Rf_length(L), REAL(out)));
    UNPROTECT(1);
    return out;
}

/* railMatrix from raw BigWig data sections (R/railMatrix.R:271-289): `sections` = list of raw vectors (per sample: the
 * inflated data sections of the leaf blocks that overlap the regions, back to back), `sec_off` = list of numeric
 * vectors (byte offset of every section, plus the total as last entry), `chrom_id` = the id each file gives the
 * chromosome.  Returns a coverage handle of sum(width(regions)) rows for dfbR_view_sums. */
SEXP dfbR_cov_from_bigwig(SEXP ctx, SEXP sections, SEXP sec_off, SEXP chrom_id, SEXP start, SEXP end, SEXP mapped_per_xm) {
    int n = Rf_length(sections);
    if (Rf_length(sec_off) != n || Rf_length(chrom_id) != n) Rf_error("one entry per sample file is needed");
    const void** sp = (const void**)R_alloc((size_t)(n > 0 ? n : 1), sizeof(void*));
    const int64_t** op = (const int64_t**)R_alloc((size_t)(n > 0 ? n : 1), sizeof(void*));
    int64_t* ns = (int64_t*)R_alloc((size_t)(n > 0 ? n : 1), sizeof(int64_t));
    uint32_t* ids = (uint32_t*)R_alloc((size_t)(n > 0 ? n : 1), sizeof(uint32_t));
    for (int s = 0; s < n; ++s) {
        SEXP o = VECTOR_ELT(sec_off, s);
        R_xlen_t k = XLENGTH(o);
        if (k < 1) Rf_error("sec_off[[%d]] needs at least the total byte count", s + 1);
        int64_t* oo = (int64_t*)R_alloc((size_t)k, sizeof(int64_t));
        for (R_xlen_t i = 0; i < k; ++i) oo[i] = (int64_t)REAL(o)[i];
        sp[s] = RAW(VECTOR_ELT(sections, s));
        op[s] = oo;
        ns[s] = (int64_t)k - 1;
        ids[s] = (uint32_t)INTEGER(chrom_id)[s];
    }
    dfb_cov* cov = NULL;
    CHECK(dfb_cov_from_bigwig((dfb_ctx*)unwrap(ctx, "railMatrix"), n, sp, op, ns, ids, (const int32_t*)INTEGER(start),
                              (const int32_t*)INTEGER(end), XLENGTH(start), Rf_isNull(mapped_per_xm) ? NULL : REAL(mapped_per_xm),
                              &cov));
    return wrap(cov, "dfb_cov", cov_finalizer, ctx);
}

/* ---------------------------------------------------------------- mergeResults -------------- */

/* regions / nulls: lists of external pointers (NULL entries: chromosome without observed regions) */
SEXP dfbR_merge_results(SEXP ctx, SEXP regions, SEXP nulls, SEXP nperm, SEXP cutoff_used, SEXP merge_area) {
    dfb_ctx* cx = (dfb_ctx*)unwrap(ctx, "mergeResults");
    int k = Rf_length(regions);
    if (Rf_length(nulls) != k) Rf_error("regions and nulls must have one entry per chromosome");
    dfb_regions** rp = (dfb_regions**)R_alloc((size_t)(k > 0 ? k : 1), sizeof(void*));
    dfb_nulls** np = (dfb_nulls**)R_alloc((size_t)(k > 0 ? k : 1), sizeof(void*));
    double** pp = (double**)R_alloc((size_t)(k > 0 ? k : 1), sizeof(double*));
    double** fp = (double**)R_alloc((size_t)(k > 0 ? k : 1), sizeof(double*));
    const char* nm[] = {"pvalues", "fwer", "totalNulls", ""};
    SEXP out = PROTECT(Rf_mkNamed(VECSXP, nm));
    SEXP pl = PROTECT(Rf_allocVector(VECSXP, k)), fl = PROTECT(Rf_allocVector(VECSXP, k));
    for (int c = 0; c < k; ++c) {
        SEXP re = VECTOR_ELT(regions, c), ne = VECTOR_ELT(nulls, c);
        rp[c] = Rf_isNull(re) ? NULL : (dfb_regions*)unwrap(re, "mergeResults");
        np[c] = Rf_isNull(ne) ? NULL : (dfb_nulls*)unwrap(ne, "mergeResults");
        R_xlen_t R = rp[c] ? (R_xlen_t)dfb_regions_count(rp[c]) : 0;
        SEXP pv = PROTECT(Rf_allocVector(REALSXP, R)), fw = PROTECT(Rf_allocVector(REALSXP, R));
        SET_VECTOR_ELT(pl, c, pv);
        SET_VECTOR_ELT(fl, c, fw);
        UNPROTECT(2);
        pp[c] = R ? REAL(pv) : NULL;
        fp[c] = R ? REAL(fw) : NULL;
    }
    int64_t total = 0;
    CHECK(dfb_merge_results(cx, k, rp, np, (int64_t)Rf_asInteger(nperm), Rf_asReal(cutoff_used), Rf_asLogical(merge_area), pp,
                            fp, &total));
    for (int c = 0; c < k; ++c) { /* NaN encodes NA (no nulls anywhere, R/calculatePvalues.R:408-419) */
        SEXP pv = VECTOR_ELT(pl, c);
        for (R_xlen_t i = 0; i < XLENGTH(pv); ++i)
            if (ISNAN(REAL(pv)[i])) REAL(pv)[i] = NA_REAL;
    }
    SET_VECTOR_ELT(out, 0, pl);
    SET_VECTOR_ELT(out, 1, fl);
    SET_VECTOR_ELT(out, 2, Rf_ScalarReal((double)total));
    UNPROTECT(3);
    return out;
}

SEXP dfbR_comm_unique_id(void) {
    SEXP id = PROTECT(Rf_allocVector(RAWSXP, DFB_COMM_ID_BYTES));
    CHECK(dfb_comm_unique_id(RAW(id)));
    UNPROTECT(1);
    return id;
}

SEXP dfbR_comm_init_rank(SEXP ctx, SEXP world, SEXP rank, SEXP id) {
    if (TYPEOF(id) != RAWSXP || XLENGTH(id) != DFB_COMM_ID_BYTES) Rf_error("communicator id must be %d raw bytes", DFB_COMM_ID_BYTES);
    CHECK(dfb_comm_init_rank((dfb_ctx*)unwrap(ctx, "communicator"), Rf_asInteger(world), Rf_asInteger(rank), RAW(id)));
    return R_NilValue;
}

/* ---------------------------------------------------------------- registration -------------- */

static const R_CallMethodDef call_methods[] = {
    {"dfbR_create", (DL_FUNC)&dfbR_create, 1},
    {"dfbR_filter", (DL_FUNC)&dfbR_filter, 9},
    {"dfbR_cov_from_rle", (DL_FUNC)&dfbR_cov_from_rle, 6},
    {"dfbR_cov_info", (DL_FUNC)&dfbR_cov_info, 1},
    {"dfbR_cov_position", (DL_FUNC)&dfbR_cov_position, 1},
    {"dfbR_cov_vector", (DL_FUNC)&dfbR_cov_vector, 3},
    {"dfbR_preprocess", (DL_FUNC)&dfbR_preprocess, 5},
    {"dfbR_chunk_lengths", (DL_FUNC)&dfbR_chunk_lengths, 3},
    {"dfbR_fstats", (DL_FUNC)&dfbR_fstats, 5},
    {"dfbR_find_regions", (DL_FUNC)&dfbR_find_regions, 9},
    {"dfbR_calculate_pvalues", (DL_FUNC)&dfbR_calculate_pvalues, 14},
    {"dfbR_merge_results", (DL_FUNC)&dfbR_merge_results, 6},
    {"dfbR_comm_unique_id", (DL_FUNC)&dfbR_comm_unique_id, 0},
    {"dfbR_comm_init_rank", (DL_FUNC)&dfbR_comm_init_rank, 4},
    {"dfbR_calc_pval", (DL_FUNC)&dfbR_calc_pval, 3},
    {"dfbR_calc_fwer", (DL_FUNC)&dfbR_calc_fwer, 6},
    {"dfbR_quantile", (DL_FUNC)&dfbR_quantile, 3},
    {"dfbR_region_matrix", (DL_FUNC)&dfbR_region_matrix, 5},
    {"dfbR_view_sums", (DL_FUNC)&dfbR_view_sums, 5},
    {"dfbR_cov_from_bigwig", (DL_FUNC)&dfbR_cov_from_bigwig, 7},
    {NULL, NULL, 0}};

void R_init_derfinderB200(DllInfo* dll) {
    R_registerRoutines(dll, NULL, call_methods, NULL, NULL);
    R_useDynamicSymbols(dll, FALSE);
}
