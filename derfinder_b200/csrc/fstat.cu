// K2 -- F-statistic scan, exact FP64 kernel.
//
// Replaces derfinderHelper::fstats.apply as called from R/calculateStats.R:123-127 and, once per
// permutation, from R/calculatePvalues.R:325-329.  Definition (tests/testthat/test_Fstats.R:52-63):
//     F = ((rss0 - rss1)/(p - p0)) / (adjustF + rss1/(n - p)),  rss = rowSums((Y %*% P)^2)
// Engine formulation: with Q = [Q0 | Qd] an orthonormal basis of span(mod) whose first p0 columns
// span mod0, and yc the (row-mean centred, legal when 1 is in span(mod0)) log2 coverage of a base,
//     S = yc' Q,  s0 = |S[0:p0]|^2,  sd = |S[p0:p]|^2,  rss1 = |yc|^2 - s0 - sd,  rss0 - rss1 = sd.
// A permutation of the sample labels (mod[idx,], mod0[idx,]) permutes the rows of Q, so all B
// null scans share one Q and differ only by a row gather: S_b = sum_k yc[k] * Q[idx_b[k], :].
//
// Arithmetic contract (bit-exact with oracle/fstat_oracle.c):
//     m  = ((y0 + y1) + ...)/n  (0 when not centred);   yc[k] = y[k] - m
//     yy = fma-chain over k of yc[k]^2;                 S[j] = fma-chain over k of yc[k]*Q[idx[k]][j]
//     s0 = fma-chain over j <  p0 of S[j]^2;            sd = fma-chain over j >= p0 of S[j]^2
//     rss1 = max((yy - s0) - sd, 0);  F = (sd/(p-p0)) / (adjustF + rss1/(n-p))
//
// Thread layout: one CTA owns a tile of TB = NT*BM consecutive bases for ALL permutations of the
// launch; the centred tile lives in shared memory (read from HBM exactly once), permutations are
// processed in groups of BP, each thread accumulating BM x (BP*PC) FP64 dot products.  In mask
// mode the epilogue thresholds in registers and emits one bit per (base, permutation) plus the
// exceedance values, compacted per (permutation, tile) -- null F-stats never reach HBM.
#include <cuda_fp16.h>
#include <math.h>
#include <type_traits>

#include <algorithm>

#include "data.cuh"

namespace dfb {

// ------------------------------------------------------------------------------------------
// host: nested orthonormal basis by modified Gram-Schmidt with re-orthogonalisation

static double dot(const double* a, const double* b, int n) {
    double s = 0;
    for (int i = 0; i < n; ++i) s += a[i] * b[i];
    return s;
}

int build_model_basis(const double* mod, int n, int p, const double* mod0, int p0, ModelBasis* out) {
    DFB_REQUIRE(mod && mod0, "models$mod / models$mod0 must not be NULL");
    DFB_REQUIRE(n > 0 && p > 0 && p0 >= 0 && p0 < p, "need 0 <= ncol(mod0) < ncol(mod) (got p0=%d, p=%d)", p0, p);
    DFB_REQUIRE(p <= n, "ncol(mod)=%d exceeds the number of samples n=%d", p, n);
    std::vector<std::vector<double>> basis;  // accepted orthonormal columns
    auto orth = [&](std::vector<double>& v) {
        for (int pass = 0; pass < 2; ++pass)
            for (auto& q : basis) {
                double c = dot(q.data(), v.data(), n);
                for (int i = 0; i < n; ++i) v[i] -= c * q[i];
            }
    };
    const double tol = 1e-9;
    for (int j = 0; j < p0; ++j) {
        std::vector<double> v(mod0 + (size_t)j * n, mod0 + (size_t)(j + 1) * n);
        double nrm0 = sqrt(dot(v.data(), v.data(), n));
        orth(v);
        double nrm = sqrt(dot(v.data(), v.data(), n));
        DFB_REQUIRE(nrm0 > 0 && nrm > tol * nrm0,
                    "models$mod0 is rank deficient (column %d); R's solve(t(mod0) %%*%% mod0) would fail", j + 1);
        for (auto& x : v) x /= nrm;
        basis.push_back(std::move(v));
    }
    for (int j = 0; j < p && (int)basis.size() < p; ++j) {
        std::vector<double> v(mod + (size_t)j * n, mod + (size_t)(j + 1) * n);
        double nrm0 = sqrt(dot(v.data(), v.data(), n));
        if (nrm0 == 0) continue;
        orth(v);
        double nrm = sqrt(dot(v.data(), v.data(), n));
        if (nrm <= tol * nrm0) continue;  // already in the span (e.g. a column shared with mod0)
        for (auto& x : v) x /= nrm;
        basis.push_back(std::move(v));
    }
    // rank of mod on its own (R's solve(t(mod) %*% mod) fails when deficient) and nestedness:
    // every column of mod0 must lie in span(mod)
    {
        std::vector<std::vector<double>> qm;
        auto orth_m = [&](std::vector<double>& v) {
            for (int pass = 0; pass < 2; ++pass)
                for (auto& q : qm) {
                    double c = dot(q.data(), v.data(), n);
                    for (int i = 0; i < n; ++i) v[i] -= c * q[i];
                }
        };
        for (int j = 0; j < p; ++j) {
            std::vector<double> v(mod + (size_t)j * n, mod + (size_t)(j + 1) * n);
            double nrm0 = sqrt(dot(v.data(), v.data(), n));
            orth_m(v);
            double nrm = sqrt(dot(v.data(), v.data(), n));
            DFB_REQUIRE(nrm0 > 0 && nrm > tol * nrm0,
                        "models$mod is rank deficient (column %d); R's solve(t(mod) %%*%% mod) would fail", j + 1);
            for (auto& x : v) x /= nrm;
            qm.push_back(std::move(v));
        }
        for (int j = 0; j < p0; ++j) {
            std::vector<double> v(mod0 + (size_t)j * n, mod0 + (size_t)(j + 1) * n);
            double nrm0 = sqrt(dot(v.data(), v.data(), n));
            orth_m(v);
            double nrm = sqrt(dot(v.data(), v.data(), n));
            DFB_REQUIRE(nrm <= 1e-7 * nrm0, "models$mod0 is not nested in models$mod (column %d of mod0 is outside "
                        "span(mod)); the engine's F-statistic needs nested models", j + 1);
        }
    }
    DFB_REQUIRE((int)basis.size() == p, "internal error: nested basis has %d columns, expected %d", (int)basis.size(),
                p);
    out->n = n;
    out->p = p;
    out->p0 = p0;
    out->q.assign((size_t)n * p, 0.0);
    for (int j = 0; j < p; ++j)
        for (int i = 0; i < n; ++i) out->q[(size_t)i * p + j] = basis[j][i];
    // centring is exact iff the constant vector is in span(mod0)
    out->centered = false;
    if (p0 > 0) {
        std::vector<double> one(n, 1.0);
        for (int pass = 0; pass < 2; ++pass)
            for (int j = 0; j < p0; ++j) {
                double c = dot(basis[j].data(), one.data(), n);
                for (int i = 0; i < n; ++i) one[i] -= c * basis[j][i];
            }
        out->centered = sqrt(dot(one.data(), one.data(), n)) <= 1e-10 * sqrt((double)n);
    }
    // column chunking for the kernel: PC columns per pass
    out->pc = p <= 6 ? p : 4;
    if (out->pc < 2) out->pc = 2;
    out->chunks = (p + out->pc - 1) / out->pc;
    out->pp = out->chunks * out->pc;
    out->qpad.assign((size_t)n * out->pp, 0.0);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < p; ++j) out->qpad[(size_t)i * out->pp + j] = out->q[(size_t)i * p + j];
    return DFB_OK;
}

// ------------------------------------------------------------------------------------------

struct FstatParams {
    const double* Y;
    int64_t ld, N;
    int n;
    const double* qpad;  // [n][pp]
    int pp, chunks, p0, p;
    const int32_t* idx;  // [B][n] row gather per permutation, or nullptr (identity)
    int64_t B;
    int groups_per_slice;  // permutation groups handled by one blockIdx.y slice
    int center;
    double adjust_f, dfnum, dfden, inv_dfnum, inv_dfden;
    int mode;       // 0: dense F out, 1: mask + compacted values
    double* fout;   // [B][N]
    double U, L;
    int fastpath;   // U > 0 && L < 0: cheap multiplicative rejection test is valid
    int two_sided;  // also emit the "dn" rows (x <= L)
    uint32_t* mask;
    int64_t W;
    int64_t* off;
    int64_t tiles;
    double* vals;
    unsigned long long capacity;
    unsigned long long* cursor;
    // observed scan only: operand cache of the tensor-core null scan (fstat_tc2.cu), written on the way --
    // centred coverage as fp16 rows [N][a16_ka] (zero padded), |yc|^2 and the row mean; null when not wanted
    __half* a16;
    int a16_ka;
    double* a16_yy;
    double* a16_m;
    // observed scan only: the row mean of the contract as dfb_preprocess computed it (same chain of additions),
    // and -- when the log2 matrix is not materialised -- the integer coverage plus the log2 table
    const double* ym;
    const int32_t* xi;
    const double* lut;
};

template <int PC, int BP, int BM, int NT>
__global__ void __launch_bounds__(NT) fstat_f64_kernel(const FstatParams P) {
    constexpr int TB = NT * BM;       // bases per tile
    constexpr int NW = TB / 32;       // mask words per tile row
    constexpr int QW = BP * PC;       // q values per sample row of the group tile
    constexpr int WPB = NT / 32;      // warps per block
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = P.n;
    double* ys = reinterpret_cast<double*>(smem_raw);  // [n][TB] centred coverage
    double* Qs = ys + (size_t)n * TB;                  // [n][pp]
    double* qs = Qs + (((size_t)n * P.pp + 1) & ~(size_t)1);  // [n][QW] gathered rows of the group (16 B aligned)
    int64_t* pre_s = reinterpret_cast<int64_t*>(qs + (size_t)n * QW);  // [2*BP*NW] value offsets
    int* cnt_s = reinterpret_cast<int*>(pre_s + 2 * BP * NW);          // [2*BP*NW] popcounts

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int64_t tile = blockIdx.x;
    const int64_t base0 = tile * TB;

    // ---- stage Q and the centred coverage tile
    for (int e = tid; e < n * P.pp; e += NT) Qs[e] = P.qpad[e];
    double yy[BM];
#pragma unroll
    for (int bm = 0; bm < BM; ++bm) {
        const int col = tid + bm * NT;
        const int64_t base = base0 + col;
        double acc = 0.0;
        if (base < P.N) {
            acc = P.Y[base];
            ys[col] = acc;
            for (int k = 1; k < n; ++k) {
                double v = P.Y[(size_t)k * P.ld + base];
                ys[(size_t)k * TB + col] = v;
                acc = __dadd_rn(acc, v);
            }
        } else {
            for (int k = 0; k < n; ++k) ys[(size_t)k * TB + col] = 0.0;
        }
        const double m = P.center ? __ddiv_rn(acc, (double)n) : 0.0;
        double s = 0.0;
        for (int k = 0; k < n; ++k) {
            double v = __dsub_rn(ys[(size_t)k * TB + col], m);
            ys[(size_t)k * TB + col] = v;
            s = fma(v, v, s);
        }
        yy[bm] = s;
    }

    const int64_t ngroups = (P.B + BP - 1) / BP;
    const int64_t g_lo = (int64_t)blockIdx.y * P.groups_per_slice;
    const int64_t g_hi = min(ngroups, g_lo + (int64_t)P.groups_per_slice);

    for (int64_t g = g_lo; g < g_hi; ++g) {
        const int64_t b0 = g * BP;
        double s0[BM][BP], sd[BM][BP];
#pragma unroll
        for (int bm = 0; bm < BM; ++bm)
#pragma unroll
            for (int bp = 0; bp < BP; ++bp) s0[bm][bp] = sd[bm][bp] = 0.0;

        for (int c = 0; c < P.chunks; ++c) {
            __syncthreads();  // previous users of qs (and, first time, Qs/ys writers) are done
            for (int e = tid; e < BP * n; e += NT) {
                const int bp = e / n, k = e - bp * n;
                const int64_t b = b0 + bp;
                double* dst = qs + (size_t)k * QW + bp * PC;
                if (b < P.B) {
                    const int row = P.idx ? P.idx[b * n + k] : k;
                    const double* src = Qs + (size_t)row * P.pp + c * PC;
#pragma unroll
                    for (int j = 0; j < PC; ++j) dst[j] = src[j];
                } else {
#pragma unroll
                    for (int j = 0; j < PC; ++j) dst[j] = 0.0;
                }
            }
            __syncthreads();

            double acc[BM][QW];
#pragma unroll
            for (int bm = 0; bm < BM; ++bm)
#pragma unroll
                for (int e = 0; e < QW; ++e) acc[bm][e] = 0.0;
#pragma unroll 2
            for (int k = 0; k < n; ++k) {
                double yv[BM];
#pragma unroll
                for (int bm = 0; bm < BM; ++bm) yv[bm] = ys[(size_t)k * TB + tid + bm * NT];
                const double2* qrow = reinterpret_cast<const double2*>(qs + (size_t)k * QW);
#pragma unroll
                for (int e2 = 0; e2 < QW / 2; ++e2) {
                    const double2 q = qrow[e2];  // broadcast LDS.128
#pragma unroll
                    for (int bm = 0; bm < BM; ++bm) {
                        acc[bm][2 * e2] = fma(yv[bm], q.x, acc[bm][2 * e2]);
                        acc[bm][2 * e2 + 1] = fma(yv[bm], q.y, acc[bm][2 * e2 + 1]);
                    }
                }
            }
#pragma unroll
            for (int bp = 0; bp < BP; ++bp)
#pragma unroll
                for (int j = 0; j < PC; ++j) {
                    const bool in0 = (c * PC + j) < P.p0;
#pragma unroll
                    for (int bm = 0; bm < BM; ++bm) {
                        const double s = acc[bm][bp * PC + j];
                        if (in0) s0[bm][bp] = fma(s, s, s0[bm][bp]);
                        else sd[bm][bp] = fma(s, s, sd[bm][bp]);
                    }
                }
        }

        // ---- epilogue
        if (P.mode == 0) {
#pragma unroll
            for (int bp = 0; bp < BP; ++bp) {
                const int64_t b = b0 + bp;
                if (b >= P.B) continue;
#pragma unroll
                for (int bm = 0; bm < BM; ++bm) {
                    const int64_t base = base0 + tid + bm * NT;
                    if (base >= P.N) continue;
                    double rss1 = __dsub_rn(__dsub_rn(yy[bm], s0[bm][bp]), sd[bm][bp]);
                    if (rss1 < 0.0) rss1 = 0.0;
                    const double num = __ddiv_rn(sd[bm][bp], P.dfnum);
                    const double den = __dadd_rn(P.adjust_f, __ddiv_rn(rss1, P.dfden));
                    P.fout[b * P.N + base] = __ddiv_rn(num, den);
                }
            }
            continue;
        }

        double fv[BM][BP];
        unsigned pass_up[BM], pass_dn[BM];  // bit bp set when (bm, bp) passes
#pragma unroll
        for (int bm = 0; bm < BM; ++bm) {
            pass_up[bm] = pass_dn[bm] = 0u;
            const int64_t base = base0 + tid + bm * NT;
#pragma unroll
            for (int bp = 0; bp < BP; ++bp) {
                fv[bm][bp] = 0.0;
                if (base >= P.N || b0 + bp >= P.B) continue;
                double rss1 = __dsub_rn(__dsub_rn(yy[bm], s0[bm][bp]), sd[bm][bp]);
                if (rss1 < 0.0) rss1 = 0.0;
                if (P.fastpath) {
                    // F < U is certain when sd/dfnum < U*(adjustF + rss1/dfden) with 1e-12 slack
                    const double lhs = sd[bm][bp] * P.inv_dfnum;
                    const double rhs = P.U * (P.adjust_f + rss1 * P.inv_dfden);
                    if (lhs < rhs * (1.0 - 1e-12)) continue;
                }
                const double num = __ddiv_rn(sd[bm][bp], P.dfnum);
                const double den = __dadd_rn(P.adjust_f, __ddiv_rn(rss1, P.dfden));
                const double f = __ddiv_rn(num, den);
                fv[bm][bp] = f;
                if (f >= P.U) pass_up[bm] |= 1u << bp;
                if (P.two_sided && f <= P.L) pass_dn[bm] |= 1u << bp;
            }
        }
        const int sides = P.two_sided ? 2 : 1;
        const int rows_g = BP * sides;  // mask rows produced by this group
        unsigned words[2][BP][BM];
#pragma unroll
        for (int sd_i = 0; sd_i < 2; ++sd_i) {
            if (sd_i < sides)
#pragma unroll
            for (int bp = 0; bp < BP; ++bp)
#pragma unroll
                for (int bm = 0; bm < BM; ++bm) {
                    const unsigned flag = ((sd_i == 0 ? pass_up[bm] : pass_dn[bm]) >> bp) & 1u;
                    const unsigned w = __ballot_sync(0xffffffffu, flag);
                    words[sd_i][bp][bm] = w;
                    if (lane == 0) {
                        const int r = bp * sides + sd_i;
                        cnt_s[r * NW + bm * WPB + warp] = __popc(w);
                        const int64_t b = b0 + bp;
                        const int64_t gw = tile * NW + bm * WPB + warp;
                        if (b < P.B && gw < P.W) P.mask[(b * sides + sd_i) * P.W + gw] = w;
                    }
                }
        }
        __syncthreads();
        if (warp == 0) {
            // exclusive scan of the rows_g*NW popcounts (row-major = (perm, side, base) order)
            const int E = rows_g * NW;
            const int per = (E + 31) / 32;
            int local = 0;
            for (int i = 0; i < per; ++i) {
                const int e = lane * per + i;
                if (e < E) local += cnt_s[e];
            }
            int inc = local;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += t;
            }
            const int total = __shfl_sync(0xffffffffu, inc, 31);
            unsigned long long gbase = 0;
            if (lane == 0 && total > 0) gbase = atomicAdd(P.cursor, (unsigned long long)total);
            gbase = __shfl_sync(0xffffffffu, gbase, 0);
            int64_t run = (int64_t)gbase + inc - local;
            for (int i = 0; i < per; ++i) {
                const int e = lane * per + i;
                if (e < E) {
                    pre_s[e] = run;
                    if (e % NW == 0) {
                        const int r = e / NW;
                        const int64_t b = b0 + r / sides;
                        if (b < P.B) P.off[(b * sides + (r % sides)) * P.tiles + tile] = run;
                    }
                    run += cnt_s[e];
                }
            }
        }
        __syncthreads();
#pragma unroll
        for (int sd_i = 0; sd_i < 2; ++sd_i) {
            if (sd_i < sides)
#pragma unroll
            for (int bp = 0; bp < BP; ++bp)
#pragma unroll
                for (int bm = 0; bm < BM; ++bm) {
                    const unsigned flag = ((sd_i == 0 ? pass_up[bm] : pass_dn[bm]) >> bp) & 1u;
                    if (flag) {
                        const int r = bp * sides + sd_i;
                        const unsigned w = words[sd_i][bp][bm];
                        const unsigned long long dst =
                            (unsigned long long)pre_s[r * NW + bm * WPB + warp] + __popc(w & ((1u << lane) - 1u));
                        if (dst < P.capacity) P.vals[dst] = fv[bm][bp];
                    }
                }
        }
        // pre_s / cnt_s are rewritten only after the next group's two __syncthreads()
    }
}

template <int PC, int BP, int BM, int NT>
static int launch_cfg(dfb_ctx* ctx, const FstatParams& P0, int kernel_family) {
    FstatParams P = P0;
    constexpr int TB = NT * BM, NW = TB / 32, QW = BP * PC;
    static_assert(QW % 2 == 0, "group tile rows are read as double2");
    size_t smem = ((size_t)P.n * TB + (((size_t)P.n * P.pp + 1) & ~(size_t)1) + (size_t)P.n * QW) * sizeof(double) +
                  (size_t)2 * BP * NW * (sizeof(int64_t) + sizeof(int));
    auto kern = fstat_f64_kernel<PC, BP, BM, NT>;
    DFB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t tiles = ceil_div(P.N, TB);
    int64_t ngroups = ceil_div(P.B, BP);
    // when the tile count cannot fill the GPU, split the permutation groups over blockIdx.y
    int64_t want = (int64_t)ctx->sm_count * 4;
    int64_t slices = tiles >= want ? 1 : std::min<int64_t>(ngroups, ceil_div(want, tiles));
    if (slices > 65535) slices = 65535;
    P.groups_per_slice = (int)ceil_div(ngroups, slices);
    slices = ceil_div(ngroups, P.groups_per_slice);
    dim3 grid((unsigned)tiles, (unsigned)slices);
    {
        KTimer t(ctx, kernel_family, 1);
        kern<<<grid, NT, smem, ctx->stream>>>(P);
    }
    DFB_CUDA(cudaGetLastError());
    return DFB_OK;
}

// tile geometry by sample count: the centred tile (n x TB doubles) must fit in shared memory
template <int PC, int BP>
static int launch_by_n(dfb_ctx* ctx, const FstatParams& P, int fam) {
    const size_t budget = 96 * 1024;
    if ((size_t)P.n * 512 * 8 <= budget) return launch_cfg<PC, BP, 4, 128>(ctx, P, fam);
    if ((size_t)P.n * 256 * 8 <= budget) return launch_cfg<PC, BP, 4, 64>(ctx, P, fam);
    if ((size_t)P.n * 128 * 8 <= (size_t)200 * 1024) return launch_cfg<PC, BP, 2, 64>(ctx, P, fam);
    set_error("F-stat kernel: %d samples exceed the shared-memory tile (max 200 samples)", P.n);
    return DFB_EINVAL;
}

int fstat_tile_bases(int n) {
    const size_t budget = 96 * 1024;
    if ((size_t)n * 512 * 8 <= budget) return 512;
    if ((size_t)n * 256 * 8 <= budget) return 256;
    return 128;
}

static int launch_fstat(dfb_ctx* ctx, const FstatParams& P, int pc, int fam) {
    switch (pc) {
        case 2: return launch_by_n<2, 6>(ctx, P, fam);
        case 3: return launch_by_n<3, 4>(ctx, P, fam);
        case 4: return launch_by_n<4, 3>(ctx, P, fam);
        case 5: return launch_by_n<5, 2>(ctx, P, fam);
        case 6: return launch_by_n<6, 2>(ctx, P, fam);
    }
    set_error("unsupported column chunk %d", pc);
    return DFB_EINVAL;
}

static void fill_common(FstatParams& P, const dfb_cov* cov, const ModelBasis& mb, const double* qpad_dev,
                        double adjust_f) {
    P.Y = cov->y.p;
    P.ld = cov->ld;
    P.N = cov->N;
    P.n = cov->n;
    P.qpad = qpad_dev;
    P.pp = mb.pp;
    P.chunks = mb.chunks;
    P.p0 = mb.p0;
    P.p = mb.p;
    P.center = mb.centered ? 1 : 0;
    P.adjust_f = adjust_f;
    P.dfnum = (double)(mb.p - mb.p0);
    P.dfden = (double)(cov->n - mb.p);
    P.inv_dfnum = 1.0 / P.dfnum;
    P.inv_dfden = 1.0 / P.dfden;
}

// Observed statistics (identity labelling): one thread per base, the coverage is streamed ONCE (coalesced
// along bases for every sample) -- the row mean of the contract comes from dfb_preprocess (same chain of
// additions), so there is no second pass -- and Q is broadcast from shared memory.  HBM-bound: 8*n bytes read
// per base from the log2 matrix, or 4*n from the integer coverage when the log2 matrix is not materialised
// (SRCI: values go through the log2 table, whose head sits in shared memory).  Same arithmetic contract as the
// tiled kernel.
template <int PC, bool EMIT, bool SRCI>
__global__ void __launch_bounds__(256, 2) fstat_observed_kernel(const FstatParams P) {
    constexpr int QS = (PC + 1) & ~1;  // row stride of a column chunk in shared memory: 16-byte aligned rows
    extern __shared__ __align__(16) double Qs_obs[];  // [chunks][n][QS], then the table head [LUT_SMEM] (SRCI)
    const int n = P.n;
    for (int e = threadIdx.x; e < P.chunks * n * QS; e += blockDim.x) {
        const int c = e / (n * QS), r = e - c * n * QS, k = r / QS, j = r - k * QS;
        Qs_obs[e] = j < PC ? P.qpad[(size_t)k * P.pp + c * PC + j] : 0.0;
    }
    double* lut_s = Qs_obs + (size_t)P.chunks * n * QS;
    if (SRCI)
        for (int e = threadIdx.x; e < LUT_SMEM; e += blockDim.x) lut_s[e] = P.lut[e];
    __syncthreads();
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; base < P.N;
         base += (int64_t)gridDim.x * blockDim.x) {
        const double m = P.center ? P.ym[base] : 0.0;
        double yy = 0.0, s0 = 0.0, sd = 0.0;
        for (int c = 0; c < P.chunks; ++c) {
            double S[PC];
#pragma unroll
            for (int j = 0; j < PC; ++j) S[j] = 0.0;
            const double* Qc = Qs_obs + (size_t)c * n * QS;
            uint4* arow = EMIT ? reinterpret_cast<uint4*>(P.a16 + (size_t)base * P.a16_ka) : nullptr;
            // Sixteen samples per step: the sixteen loads are issued back to back (no predicate per load -- with
            // predicates the compiler sinks every load next to its table look-up and the warp waits for each of
            // them in turn), then the table look-ups, then the products.  When the fp16 operand is emitted a step
            // writes one whole 32-byte sector of the base's row (two 16-byte stores; half sectors would make L2
            // read the other half back from DRAM).  FULL: all sixteen samples exist; FIRST: first column chunk
            // (|yc|^2 and the fp16 row are produced there).
            auto step = [&](const int k0, auto full_tag, auto first_tag) {
                constexpr bool FULL = decltype(full_tag)::value, FIRST = decltype(first_tag)::value;
                double yv[16];
                if (SRCI) {
                    int32_t raw[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e)
                        raw[e] = __ldg(P.xi + (size_t)(FULL ? k0 + e : min(k0 + e, n - 1)) * P.ld + base);
                    asm volatile("" ::: "memory");
                    int mx = raw[0];
#pragma unroll
                    for (int e = 1; e < 16; ++e) mx = max(mx, raw[e]);
                    if (mx < LUT_SMEM) {
#pragma unroll
                        for (int e = 0; e < 16; ++e) yv[e] = lut_s[raw[e]];
                    } else {  // rare: a value beyond the table head
#pragma unroll
                        for (int e = 0; e < 16; ++e) yv[e] = raw[e] < LUT_SMEM ? lut_s[raw[e]] : __ldg(P.lut + raw[e]);
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < 16; ++e)
                        yv[e] = __ldg(P.Y + (size_t)(FULL ? k0 + e : min(k0 + e, n - 1)) * P.ld + base);
                    asm volatile("" ::: "memory");
                }
                __half2 h2[8];
                const double* qk = Qc + (size_t)k0 * QS;
#pragma unroll
                for (int e = 0; e < 16; e += 2) {
                    double v2[2];
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        v2[u] = 0.0;
                        if (FULL || k0 + e + u < n) {
                            const double v = __dsub_rn(yv[e + u], m);
                            if (FIRST) yy = fma(v, v, yy);
                            const double* q = qk + (e + u) * QS;
#pragma unroll
                            for (int j = 0; j < PC; ++j) S[j] = fma(v, q[j], S[j]);
                            v2[u] = v;
                        }
                    }
                    if (EMIT && FIRST) h2[e >> 1] = __halves2half2(__double2half(v2[0]), __double2half(v2[1]));
                }
                if (EMIT && FIRST) {
                    arow[k0 >> 3] = reinterpret_cast<const uint4*>(h2)[0];
                    arow[(k0 >> 3) + 1] = reinterpret_cast<const uint4*>(h2)[1];
                }
            };
            int k0 = 0;
            if (c == 0) {
                for (; k0 + 16 <= n; k0 += 16) step(k0, std::true_type(), std::true_type());
                if (k0 < n) step(k0, std::false_type(), std::true_type());
                if (EMIT)
                    for (int kz = (n + 15) & ~15; kz < P.a16_ka; kz += 8) arow[kz >> 3] = make_uint4(0u, 0u, 0u, 0u);
            } else {
                for (; k0 + 16 <= n; k0 += 16) step(k0, std::true_type(), std::false_type());
                if (k0 < n) step(k0, std::false_type(), std::false_type());
            }
#pragma unroll
            for (int j = 0; j < PC; ++j) {
                if (c * PC + j < P.p0) s0 = fma(S[j], S[j], s0);
                else sd = fma(S[j], S[j], sd);
            }
        }
        if (EMIT) {
            P.a16_yy[base] = yy;
            P.a16_m[base] = m;
        }
        double rss1 = __dsub_rn(__dsub_rn(yy, s0), sd);
        if (rss1 < 0.0) rss1 = 0.0;
        const double num = __ddiv_rn(sd, P.dfnum);
        const double den = __dadd_rn(P.adjust_f, __ddiv_rn(rss1, P.dfden));
        P.fout[base] = __ddiv_rn(num, den);
    }
}

int fstat_observed(dfb_ctx* ctx, const dfb_cov* cov, const ModelBasis& mb, double adjust_f, double* f_dev) {
    DevBuf<double> qd;
    DFB_TRY(qd.alloc(ctx, mb.qpad.size()));
    DFB_TRY(h2d(ctx, qd.p, mb.qpad.data(), mb.qpad.size() * sizeof(double)));
    FstatParams P;
    memset(&P, 0, sizeof P);
    fill_common(P, cov, mb, qd.p, adjust_f);
    P.idx = nullptr;
    P.B = 1;
    P.mode = 0;
    P.fout = f_dev;
    // The tensor-core null scan of calculatePvalues() needs the centred coverage as fp16 rows plus |yc|^2 and
    // the row mean: exactly what this kernel has in registers.  Writing them here costs 2 bytes per element
    // (the stand-alone builder re-reads the whole FP64 matrix: 3.9 ms at the chr1 shape).
    {
        dfb_cov* cw = const_cast<dfb_cov*>(cov);
        cw->a16_center = -1;
        if (ctx->opt_a16_fused && ctx->opt_screen && null2_wanted(ctx, cov->n) && cov->n <= 255 && mb.p >= 2 && mb.p <= 8 &&
            cov->n - mb.p > 0) {
            const int KA = (int)ceil_div(round_up(cov->n, 16), 64) * 64;
            const int64_t rows_pad = ceil_div(cov->N, 128) * 128;
            DFB_TRY(cw->a16.alloc(ctx, (size_t)rows_pad * KA));
            DFB_TRY(cw->a16_yy.alloc(ctx, (size_t)std::max<int64_t>(cov->N, 1)));
            DFB_TRY(cw->a16_m.alloc(ctx, (size_t)std::max<int64_t>(cov->N, 1)));
            if (rows_pad > cov->N)  // rows of the last 128-base tile beyond N
                DFB_CUDA(cudaMemsetAsync(cw->a16.p + (size_t)cov->N * KA, 0, (size_t)(rows_pad - cov->N) * KA * 2, ctx->stream));
            P.a16 = reinterpret_cast<__half*>(cw->a16.p);
            P.a16_ka = KA;
            P.a16_yy = cw->a16_yy.p;
            P.a16_m = cw->a16_m.p;
            cw->a16_center = mb.centered ? 1 : 0;
            cw->a16_ka = KA;
            cw->a16_rows = rows_pad;
        }
    }
    DFB_REQUIRE(cov->has_ym, "internal error: the row means of the log2 matrix are missing");
    P.ym = cov->ym.p;
    const bool srci = cov->y_lazy && !cov->y.p;  // log2 matrix not materialised: integer coverage + table
    if (srci) {
        DFB_TRY(ensure_lut(ctx, cov->scalefac));
        P.xi = cov->xi.p;
        P.lut = ctx->lut_dev;
    }
    // shared memory: the basis as [chunks][n][pc rounded up to even] + the head of the log2 table
    const size_t smem = ((size_t)mb.chunks * mb.n * ((mb.pc + 1) & ~1) + (srci ? (size_t)LUT_SMEM : 0)) * sizeof(double);
    auto launch = [&](auto kern) -> int {
        if (smem > 48 * 1024) DFB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        // exactly one resident wave (grid-stride loop inside): a grid of 8 blocks per SM with 3 resident ran 2.67 waves,
        // the last one a third empty, and every block reloads the basis and the table head
        int occ = 0;
        DFB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, smem));
        const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(cov->N, 256), (int64_t)ctx->sm_count * std::max(occ, 1)));
        KTimer t(ctx, DFB_K_FSTAT, 1);
        kern<<<grid, 256, smem, ctx->stream>>>(P);
        return DFB_OK;
    };
    auto by_src = [&](auto emit_tag, auto pc_tag) -> int {
        constexpr bool E = decltype(emit_tag)::value;
        constexpr int PCV = decltype(pc_tag)::value;
        if (srci) return launch(fstat_observed_kernel<PCV, E, true>);
        return launch(fstat_observed_kernel<PCV, E, false>);
    };
    auto by_pc = [&](auto emit_tag) -> int {
        switch (mb.pc) {
            case 2: return by_src(emit_tag, std::integral_constant<int, 2>());
            case 3: return by_src(emit_tag, std::integral_constant<int, 3>());
            case 4: return by_src(emit_tag, std::integral_constant<int, 4>());
            case 5: return by_src(emit_tag, std::integral_constant<int, 5>());
            default: return by_src(emit_tag, std::integral_constant<int, 6>());
        }
    };
    if (P.a16) DFB_TRY(by_pc(std::true_type()));
    else DFB_TRY(by_pc(std::false_type()));
    DFB_CUDA(cudaGetLastError());
    // qd is released stream-ordered after the kernel
    return DFB_OK;
}

// One batch of permutations -> mask + compacted exceedances.  idx_dev: [B][n] on the device.
int fstat_perm_batch(dfb_ctx* ctx, const dfb_cov* cov, const ModelBasis& mb, const double* qpad_dev,
                     double adjust_f, const int32_t* idx_dev, int64_t B, double lo, double hi, bool two_sided,
                     size_t capacity_hint, PermBatch* out) {
    DFB_TRY(ensure_y(ctx, cov));
    const int sides = two_sided ? 2 : 1;
    const int tb = fstat_tile_bases(cov->n);
    out->rows = B * sides;
    out->W = ceil_div(cov->N, 32);
    out->tile_bases = tb;
    out->tiles = ceil_div(cov->N, tb);
    DFB_TRY(out->mask.alloc(ctx, (size_t)out->rows * out->W));
    DFB_TRY(out->off.alloc(ctx, (size_t)out->rows * out->tiles));
    DevBuf<unsigned long long> cursor;
    DFB_TRY(cursor.alloc(ctx, 1));
    size_t dense = (size_t)B * sides * (size_t)cov->N;
    size_t cap = std::min(dense, std::max<size_t>(capacity_hint, 1024));
    for (int attempt = 0; attempt < 2; ++attempt) {
        DFB_TRY(out->vals.alloc(ctx, cap));
        DFB_TRY(cursor.zero());
        FstatParams P;
        memset(&P, 0, sizeof P);
        fill_common(P, cov, mb, qpad_dev, adjust_f);
        P.idx = idx_dev;
        P.B = B;
        P.mode = 1;
        P.U = hi;
        P.L = lo;
        P.fastpath = (hi > 0.0 && lo < 0.0) ? 1 : 0;
        P.two_sided = two_sided ? 1 : 0;
        P.mask = out->mask.p;
        P.W = out->W;
        P.off = out->off.p;
        P.tiles = out->tiles;
        P.vals = out->vals.p;
        P.capacity = cap;
        P.cursor = cursor.p;
        DFB_TRY(launch_fstat(ctx, P, mb.pc, DFB_K_FSTAT));
        unsigned long long used = 0;
        DFB_TRY(d2h(ctx, &used, cursor.p, sizeof used));
        out->nvals = (int64_t)used;
        if (used <= cap) return DFB_OK;
        cap = (size_t)used;  // exact requirement is known now: one more pass
    }
    set_error("exceedance buffer overflow after resize (internal error)");
    return DFB_ECUDA;
}

}  // namespace dfb

using namespace dfb;

extern "C" int dfb_model_basis(const double* mod, int n, int p, const double* mod0, int p0, double* q,
                               int* centered) {
    ModelBasis mb;
    DFB_TRY(build_model_basis(mod, n, p, mod0, p0, &mb));
    if (q)
        for (int j = 0; j < p; ++j)
            for (int i = 0; i < n; ++i) q[(size_t)j * n + i] = mb.q[(size_t)i * p + j];
    if (centered) *centered = mb.centered ? 1 : 0;
    return DFB_OK;
}
