/*
 * oracle/fstat_oracle.c -- CPU restatement (plain C) of the F-statistic arithmetic.
 *
 * TEST INFRASTRUCTURE ONLY: linked/loaded by tests/, __graft_entry__.smoke() and bench.py's
 * CPU-baseline legs, never by the product library.
 *
 * Reference: derfinderHelper::fstats.apply as called from R/calculateStats.R:123-127 and
 * R/calculatePvalues.R:325-329 (un-vendored; Bioconductor derfinderHelper >= 1.1.0, DESCRIPTION:22),
 * whose definition the reference pins in tests/testthat/test_Fstats.R:52-63:
 *     F = ((rss0 - rss1)/(df1 - df0)) / (adjustF + rss1/(n - df1)),  rss = rowSums((Y %*% P)^2).
 *
 *   oc_fstats_projector : that definition verbatim (residual projectors, two products, row sums).
 *   oc_fstats_basis     : the algebraically identical nested-orthonormal-basis form the engine
 *                         evaluates, with the SAME operation order as the CUDA kernel
 *                         (derfinder_b200/csrc/fstat.cu), so GPU results can be compared bit for
 *                         bit.  Pinned against the golden genomeFstats through
 *                         tests/test_oracle_golden.py::test_c_oracle_*.
 *
 * Build: gcc -O2 -fPIC -shared -mfma -ffp-contract=off (see oracle/Makefile); fma() must compile to
 * a single-rounding hardware FMA.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* Y: sample-major [n][ld]; q: [n][p] row-major orthonormal basis (first p0 columns span mod0);
 * idx: n row indices of the permutation (NULL = identity); out: N doubles. */
void oc_fstats_basis(const double* Y, int64_t N, int n, int64_t ld, const double* q, int p, int p0,
                     const int32_t* idx, int center, double adjust_f, double* out) {
    double* yc = (double*)malloc(sizeof(double) * (size_t)n);
    double* S = (double*)malloc(sizeof(double) * (size_t)p);
    const double dfnum = (double)(p - p0), dfden = (double)(n - p);
    for (int64_t i = 0; i < N; ++i) {
        double acc = Y[i];
        for (int k = 1; k < n; ++k) acc = acc + Y[(size_t)k * ld + i];
        const double m = center ? acc / (double)n : 0.0;
        double yy = 0.0;
        for (int k = 0; k < n; ++k) {
            yc[k] = Y[(size_t)k * ld + i] - m;
            yy = fma(yc[k], yc[k], yy);
        }
        for (int j = 0; j < p; ++j) S[j] = 0.0;
        for (int k = 0; k < n; ++k) {
            const double* row = q + (size_t)(idx ? idx[k] : k) * p;
            for (int j = 0; j < p; ++j) S[j] = fma(yc[k], row[j], S[j]);
        }
        double s0 = 0.0, sd = 0.0;
        for (int j = 0; j < p0; ++j) s0 = fma(S[j], S[j], s0);
        for (int j = p0; j < p; ++j) sd = fma(S[j], S[j], sd);
        double rss1 = (yy - s0) - sd;
        if (rss1 < 0.0) rss1 = 0.0;
        const double num = sd / dfnum;
        const double den = adjust_f + rss1 / dfden;
        out[i] = num / den;
    }
    free(yc);
    free(S);
}

/* P1, P0: [n][n] row-major residual projectors I - M (M'M)^-1 M' (built by the caller).
 * The reference's shape of work: two dense [N x n]*[n x n] products per call. */
void oc_fstats_projector(const double* Y, int64_t N, int n, int64_t ld, const double* P1, const double* P0,
                         int p, int p0, double adjust_f, double* out) {
    double* y = (double*)malloc(sizeof(double) * (size_t)n);
    for (int64_t i = 0; i < N; ++i) {
        for (int k = 0; k < n; ++k) y[k] = Y[(size_t)k * ld + i];
        double rss1 = 0.0, rss0 = 0.0;
        for (int c = 0; c < n; ++c) {
            double r1 = 0.0, r0 = 0.0;
            for (int k = 0; k < n; ++k) {
                r1 += y[k] * P1[(size_t)k * n + c];
                r0 += y[k] * P0[(size_t)k * n + c];
            }
            rss1 += r1 * r1;
            rss0 += r0 * r0;
        }
        out[i] = ((rss0 - rss1) / (double)(p - p0)) / (adjust_f + rss1 / (double)(n - p));
    }
    free(y);
}

/* sum(Views(Rle(x), IRanges(start, end))) with IRanges' run-by-run accumulation: a run of k equal
 * adjacent values contributes value*k as one product, runs are added left to right in double
 * (R/findRegions.R:242-243, 275-276 call sum()/mean() on RleViews).  start/end 1-based inclusive. */
void oc_view_sums(const double* x, const int64_t* start, const int64_t* end, int64_t R, double* out) {
    for (int64_t r = 0; r < R; ++r) {
        int64_t i = start[r] - 1, e = end[r];
        double acc = 0.0;
        while (i < e) {
            double cur = x[i];
            int64_t j = i + 1;
            while (j < e && x[j] == cur) ++j;
            acc = acc + cur * (double)(j - i);
            i = j;
        }
        out[r] = acc;
    }
}
