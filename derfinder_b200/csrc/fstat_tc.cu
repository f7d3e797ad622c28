// K2 (fast path) -- permutation F-stat screening contraction on the 5th-gen tensor cores, followed
// by an exact FP64 refinement of the candidates.
//
// Reference work replaced: the B calls of derfinderHelper::fstats.apply inside the permutation loop
// of calculatePvalues (R/calculatePvalues.R:306-343), i.e. B full F-stat scans whose only use is
// "which bases reach the cutoff, and what is F there".
//
// Screening.  With S = yc' [Q0|Qd] (see fstat.cu) the test F >= U is algebraically
//     a*|S_d|^2 + c*|S_0|^2 >= U*(adjustF + |yc|^2/dfden),   a = 1/dfnum + U/dfden,  c = U/dfden,
// whose right-hand side does not depend on the permutation.  S for 128 bases x PS permutations is
// one [128 x K] x [K x N] product on tcgen05 (kind::f16, bf16 operands, FP32 accumulation in TMEM).
// Operands are split hi/lo (y = yh + yl, q = qh + ql, bf16 each) and three of the four cross terms
// are stacked along K ([yh|yh|yl] x [qh|ql|qh]), so |S~ - S| <= delta*|yc| with delta ~ 1e-5 (bound
// derived in DESIGN.md; the launch uses a 2x safety factor and tests measure it).  The epilogue
// reads TMEM, forms L~ = sum_j w_j S~_j^2 and compares it with a per-base threshold lowered by the
// error bound; the result is a CANDIDATE bit mask that contains every true exceedance.  Null
// F-stats are never materialised.
//
// Refinement.  Every candidate (base, permutation) is recomputed with the exact FP64 arithmetic
// contract of fstat.cu (same fma chains, bit-identical to the dense kernel and the C oracle),
// thresholded exactly, and emitted in the same "mask + compacted values" form the region finder
// consumes.  Results are therefore identical with the screen on or off.
//
// Layout: A tile = 128 bases x Kp bf16, B tile = Nmma columns x Kp bf16, both K-major in the
// 128-byte swizzled UMMA layout (rows of 64 bf16, 8-row / 1024-byte atoms, 16-byte chunk index
// XOR (row % 8)); the B tiles of all permutation chunks are laid out by a small prep kernel so a
// plain bulk copy (cp.async.bulk + mbarrier) brings a ready tile into shared memory.  The A tile
// is built once per CTA and reused for every permutation chunk.
#include <cuda_bf16.h>
#include <math.h>
#include <stdlib.h>

#include <algorithm>

#include "data.cuh"
#include "tc_ptx.cuh"

namespace dfb {

constexpr int TC_THREADS = 160;  // warps 0-3: A-tile build + epilogue, warp 4: bulk copies + MMA issue

// ------------------------------------------------------------------------------------------
// geometry shared by the prep kernel, the screen and the refinement

struct ScreenGeom {
    int n, npad, Kp, KB;  // samples, samples padded to 16, K = 3*npad padded to 64, K blocks of 64
    int p, p0, pst;       // model columns, column stride per permutation (power of two >= p)
    int nmma, ps;         // MMA N (TMEM columns per stage) and permutations per stage
    int64_t B, chunks, Bpad;
    int64_t tiles;        // 128-base tiles
    size_t smem;
};

// ------------------------------------------------------------------------------------------
// B operand: for every chunk of `ps` permutations a [nmma x Kp] bf16 tile, pre-swizzled.
// Row r = q*pst + j holds column j of permutation q; K layout [qh | ql | qh] matching [yh | yh | yl].

__global__ void __launch_bounds__(128) build_qb_kernel(const double* __restrict__ q /*[n][p] row-major*/,
                                                       const int32_t* __restrict__ idx /*[B][n]*/, ScreenGeom G,
                                                       double sw0, double swd, unsigned char* __restrict__ out) {
    // columns are pre-scaled by sqrt(w_j) (w = c for the p0 null columns, a for the others) so the
    // epilogue's test is a plain sum of squares
    // one thread per 16-byte chunk of 8 consecutive K values of one row
    const int64_t chunks_per_row = G.Kp / 8;
    const int64_t total = G.chunks * G.nmma * chunks_per_row;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = t / (G.nmma * chunks_per_row);
        const int64_t rem = t - c * G.nmma * chunks_per_row;
        const int row = (int)(rem / chunks_per_row);
        const int kc = (int)(rem - (int64_t)row * chunks_per_row);  // 8-column chunk along K
        const int qi = row / G.pst, j = row - qi * G.pst;
        const int64_t b = c * G.ps + qi;
        __nv_bfloat16 vals[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int col = kc * 8 + e;
            const int term = col / G.npad, k = col - term * G.npad;
            float f = 0.f;
            if (b < G.B && j < G.p && term < 3 && k < G.n) {
                const double v = q[(size_t)idx[b * G.n + k] * G.p + j] * (j < G.p0 ? sw0 : swd);
                const __nv_bfloat16 h = __double2bfloat16(v);
                if (term == 1) f = __bfloat162float(__double2bfloat16(v - (double)__bfloat162float(h)));
                else f = __bfloat162float(h);
            }
            vals[e] = __float2bfloat16(f);  // exact: f is already a bf16 value
        }
        const int kb = (kc * 8) / 64, colb = (kc * 8) % 64;
        unsigned char* dst = out + ((size_t)(c * G.KB + kb) * G.nmma) * 128 + sw128_offset(row, colb);
        *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(vals);
    }
}

// ------------------------------------------------------------------------------------------
// screening kernel

struct ScreenParams {
    const double* Y;
    int64_t ld, N;
    ScreenGeom G;
    const unsigned char* qb;  // [chunks][KB][nmma][128 B]
    int center;
    double U, adjust_f, dfnum, dfden, delta;
    uint32_t* cand;      // [tiles][4 warps][Bpad] candidate words
    uint32_t idesc;
};

// PST: TMEM column stride per permutation (power of two), PCOLS: columns actually summed (= p)
template <int PST, int PCOLS>
__global__ void __launch_bounds__(TC_THREADS) fstat_screen_kernel(const ScreenParams P) {
    extern __shared__ __align__(1024) unsigned char smem_tc_raw[];
    // SWIZZLE_128B atoms need 1024-byte alignment: align by hand (the launch reserves the slack)
    unsigned char* smem_tc = smem_tc_raw + ((1024u - (smem_u32(smem_tc_raw) & 1023u)) & 1023u);
    const ScreenGeom& G = P.G;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t a_bytes = (uint32_t)G.KB * TC_M * 128u;
    const uint32_t b_bytes = (uint32_t)G.KB * G.nmma * 128u;
    unsigned char* As = smem_tc;
    unsigned char* Bs = smem_tc + a_bytes;  // two stages
    uint64_t* bars = reinterpret_cast<uint64_t*>(Bs + 2 * b_bytes);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
    uint32_t* words_s = tmem_slot + 4;  // [4 warps][32] candidate words of the current chunk
    const uint32_t full_b = smem_u32(bars + 0), empty_b = smem_u32(bars + 2);
    const uint32_t tmem_full = smem_u32(bars + 4), tmem_empty = smem_u32(bars + 6);

    const int64_t tile = blockIdx.x;
    const int64_t base0 = tile * TC_M;

    if (warp == 4) {
        if (lane == 0) {
            for (int s = 0; s < 2; ++s) {
                mbar_init(full_b + 8 * s, 1);
                mbar_init(empty_b + 8 * s, 1);
            }
            for (int s = 0; s < 2; ++s) {
                mbar_init(tmem_full + 8 * s, 1);
                mbar_init(tmem_empty + 8 * s, 4);
            }
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc(smem_u32(tmem_slot), (uint32_t)(2 * G.nmma));
    }

    // ---- A tile: thread t (< 128) converts base row t
    float thr = __int_as_float(0x7f800000);  // +inf: rows beyond N never pass
    if (tid < TC_M) {
        const int64_t base = base0 + tid;
        const bool valid = base < P.N;
        double m = 0.0, yy = 0.0;
        if (valid && P.center) {
            double acc = 0.0;
            for (int k0 = 0; k0 < G.n; k0 += 8) {  // 8 independent loads in flight, then the ordered sum
                double v[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = (k0 + e < G.n) ? P.Y[(size_t)(k0 + e) * P.ld + base] : 0.0;
#pragma unroll
                for (int e = 0; e < 8; ++e)
                    if (k0 + e < G.n) acc = (k0 + e == 0) ? v[e] : __dadd_rn(acc, v[e]);
            }
            m = __ddiv_rn(acc, (double)G.n);
        }
        // every sample is loaded and split once: 8 samples -> one 16-byte chunk of yh (written to
        // the K ranges of terms 0 and 1) and one of yl (term 2); K columns [3*npad, Kp) are zero
        const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
        auto put = [&](int col, const uint4& val) {  // col: multiple of 8
            *reinterpret_cast<uint4*>(As + (size_t)(col >> 6) * TC_M * 128 + sw128_offset(tid, col & 63)) = val;
        };
        for (int k0 = 0; k0 < G.npad; k0 += 8) {
            __nv_bfloat16 hi[8], lo[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int k = k0 + e;
                double v = 0.0;
                if (valid && k < G.n) {
                    v = __dsub_rn(P.Y[(size_t)k * P.ld + base], m);
                    yy = fma(v, v, yy);
                }
                hi[e] = __double2bfloat16(v);
                lo[e] = __double2bfloat16(v - (double)__bfloat162float(hi[e]));
            }
            const uint4 h4 = *reinterpret_cast<const uint4*>(hi), l4 = *reinterpret_cast<const uint4*>(lo);
            put(k0, h4);
            put(G.npad + k0, h4);
            put(2 * G.npad + k0, l4);
        }
        for (int col = 3 * G.npad; col < G.Kp; col += 8) put(col, zero4);
        if (valid) {
            // candidate iff  a*sd~ + c*s0~ >= (sqrt(T) - kappa*delta*|yc|)^2,  T = U*(adjustF + yy/dfden)
            const double a = 1.0 / P.dfnum + P.U / P.dfden, cc = P.U / P.dfden;
            const double kappa = sqrt(a * (double)(G.p - G.p0) + cc * (double)G.p0);
            const double T = P.U * (P.adjust_f + yy / P.dfden);
            const double r = sqrt(T) - kappa * P.delta * sqrt(yy);
            const double t2 = r > 0.0 ? r * r * (1.0 - 1e-5) : 0.0;  // 1e-5: fp32 epilogue rounding
            thr = (t2 == t2) ? (float)t2 : __int_as_float(0x7fc00000);  // NaN data -> never a candidate
            if (thr > (float)t2 && thr > 0.f) thr = __int_as_float(__float_as_int(thr) - 1);  // round down
        }
    }
    fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 4) {
        // ===== producer + MMA issuer (one lane) =====
        if (lane == 0) {
            const uint32_t a_addr = smem_u32(As), b_addr = smem_u32(Bs);
            const unsigned char* src = P.qb;
            mbar_expect_tx(full_b, b_bytes);
            bulk_g2s(b_addr, src, b_bytes, full_b);
            for (int64_t c = 0; c < G.chunks; ++c) {
                const int s = (int)(c & 1);
                const uint32_t ph = (uint32_t)((c >> 1) & 1);
                if (c + 1 < G.chunks) {  // prefetch the next B tile into the other stage
                    const int s1 = (int)((c + 1) & 1);
                    const uint32_t ph1 = (uint32_t)(((c + 1) >> 1) & 1);
                    mbar_wait(empty_b + 8 * s1, ph1 ^ 1);
                    mbar_expect_tx(full_b + 8 * s1, b_bytes);
                    bulk_g2s(b_addr + s1 * b_bytes, src + (size_t)(c + 1) * b_bytes, b_bytes, full_b + 8 * s1);
                }
                mbar_wait(full_b + 8 * s, ph);
                mbar_wait(tmem_empty + 8 * s, ph ^ 1);
                tc_fence_after();
                const uint32_t d_addr = tmem_base + (uint32_t)(s * G.nmma);
                for (int kb = 0; kb < G.KB; ++kb) {
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const uint64_t ad = umma_desc_sw128(a_addr + (uint32_t)kb * TC_M * 128u + kk * 32u);
                        const uint64_t bd =
                            umma_desc_sw128(b_addr + s * b_bytes + (uint32_t)kb * G.nmma * 128u + kk * 32u);
                        umma_bf16(d_addr, ad, bd, P.idesc, (kb | kk) ? 1u : 0u);
                    }
                }
                umma_commit(empty_b + 8 * s);    // B stage reusable once these MMAs have read it
                umma_commit(tmem_full + 8 * s);  // accumulator ready for the epilogue
            }
        }
    } else {
        // ===== epilogue: thread = base row; TMEM lanes 32*warp .. 32*warp+31 =====
        uint32_t* out = P.cand + ((size_t)tile * 4 + warp) * G.Bpad;
        constexpr int per16 = 16 / PST;  // permutations per 16-column load
        for (int64_t c = 0; c < G.chunks; ++c) {
            const int s = (int)(c & 1);
            const uint32_t ph = (uint32_t)((c >> 1) & 1);
            mbar_wait(tmem_full + 8 * s, ph);
            tc_fence_after();
            uint32_t* wrow = words_s + warp * 32;
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(s * G.nmma);
            for (int g = 0; g < G.nmma / 16; ++g) {
                float v[16];
                tmem_ld16(taddr + (uint32_t)(g * 16), v);
#pragma unroll
                for (int qq = 0; qq < per16; ++qq) {
                    float acc = v[qq * PST] * v[qq * PST];
#pragma unroll
                    for (int j = 1; j < PCOLS; ++j) acc = fmaf(v[qq * PST + j], v[qq * PST + j], acc);
                    const unsigned word = __ballot_sync(0xffffffffu, acc >= thr);
                    if (lane == 0) wrow[g * per16 + qq] = word;
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty + 8 * s);
            const int64_t b = c * G.ps + lane;  // one coalesced store of the chunk's words per warp
            if (lane < G.ps && b < G.Bpad) out[b] = b < G.B ? wrow[lane] : 0u;
            __syncwarp();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem_base, (uint32_t)(2 * G.nmma));
}

// ------------------------------------------------------------------------------------------
// refinement: exact FP64 F for every candidate, exact threshold, compaction (same output format as
// the dense kernel with 128-base tiles)

struct RefineParams {
    const double* Y;
    int64_t ld, N;
    int n, p, p0;
    const double* q;      // [n][p] row-major
    const int32_t* idx;   // [B][n]
    int64_t B, Bpad;
    int center;
    double adjust_f, dfnum, dfden, U;
    const uint32_t* cand;  // [tiles][4][Bpad]
    uint32_t* mask;        // [B][W]
    int64_t W;
    int64_t* off;          // [B][W] offset of the first value of every 32-base word
    int64_t tiles128;      // 128-base staging tiles
    double* vals;
    unsigned long long capacity;
    unsigned long long* cursor;
};

constexpr int RF_THREADS = 128;
constexpr int RF_CHUNK = 2048;  // values a warp reserves from the global cursor at a time

static int refine_qstride(int p) { return (p + 1) & ~1; }  // rows of Q padded to 16 bytes

static size_t refine_smem_bytes(int n, int p) {
    // ys [n][128] + Qs [n][qstride] + yy [128] + per warp: list [1024] u16, ew [32] u32, perm rows [32][n] u8
    return ((size_t)n * TC_M + (size_t)n * refine_qstride(p) + TC_M) * 8 +
           4 * (1024 * 2 + 32 * 4 + (size_t)round_up(32 * n, 16)) + 64;
}

// Persistent, warp-autonomous refinement.  A CTA stages one 128-base tile (centred FP64 log2
// coverage, read from HBM once); each of its 4 warps then owns 32 bases (= one mask word per
// permutation) and walks the permutations 32 at a time without any block-level barrier:
//   lane = permutation: candidate word -> per-warp work list, sorted by (permutation, base);
//   lane = work item  : exact F with the FP64 contract of fstat.cu; exceedances are compacted
//                       straight into the value buffer (ballot prefix keeps the list order, which
//                       is the (permutation, base) order the region finder needs) and set a bit
//                       of the exact word;
//   lane = permutation: exact word + value offset out.
// Value storage comes from per-warp chunks of the global cursor (one atomic per RF_CHUNK values).
// Value tiles are 32 bases wide: off[b][w].
// PT: number of model columns when <= 8 (accumulators in registers), 0 = generic
template <int PT>
__global__ void __launch_bounds__(RF_THREADS) fstat_refine_kernel(const RefineParams P) {
    extern __shared__ __align__(16) unsigned char smem_rf[];
    const int n = P.n, p = P.p;
    const int qs = (p + 1) & ~1;
    double* ys = reinterpret_cast<double*>(smem_rf);  // [n][128] centred
    double* Qs = ys + (size_t)n * TC_M;               // [n][qs]
    double* yy_s = Qs + (size_t)n * qs;               // [128]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t per_warp = 1024 * 2 + 32 * 4 + (size_t)((32 * n + 15) & ~15);
    unsigned char* wbase = reinterpret_cast<unsigned char*>(yy_s + TC_M) + (size_t)warp * per_warp;
    uint16_t* list_w = reinterpret_cast<uint16_t*>(wbase);        // [1024] (perm << 5) | base
    uint32_t* ew_w = reinterpret_cast<uint32_t*>(list_w + 1024);  // [32] exact words
    unsigned char* ix_w = reinterpret_cast<unsigned char*>(ew_w + 32);  // [32][n] permutation rows

    for (int e = tid; e < n * qs; e += RF_THREADS) {
        const int k = e / qs, j = e - k * qs;
        Qs[e] = j < p ? P.q[(size_t)k * p + j] : 0.0;
    }
    unsigned long long chunk_ptr = 0;  // warp-uniform: next free value slot of the warp's chunk
    int chunk_left = 0;
    const bool num_is_sd = P.dfnum == 1.0;  // x / 1.0 == x: skip one IEEE division

    for (int64_t tile = blockIdx.x; tile < P.tiles128; tile += gridDim.x) {
        const int64_t base0 = tile * TC_M;
        __syncthreads();  // every warp is done with the previous tile (and Qs is staged)
        {
            const int64_t base = base0 + tid;
            double acc = 0.0;
            if (base < P.N) {
                // batches of 8 independent loads, then the ordered sum
                for (int k0 = 0; k0 < n; k0 += 8) {
                    double v[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) v[e] = (k0 + e < n) ? P.Y[(size_t)(k0 + e) * P.ld + base] : 0.0;
#pragma unroll
                    for (int e = 0; e < 8; ++e)
                        if (k0 + e < n) {
                            ys[(size_t)(k0 + e) * TC_M + tid] = v[e];
                            acc = (k0 + e == 0) ? v[e] : __dadd_rn(acc, v[e]);
                        }
                }
            } else {
                for (int k = 0; k < n; ++k) ys[(size_t)k * TC_M + tid] = 0.0;
            }
            const double m = P.center ? __ddiv_rn(acc, (double)n) : 0.0;
            double s = 0.0;
            for (int k = 0; k < n; ++k) {
                const double v = __dsub_rn(ys[(size_t)k * TC_M + tid], m);
                ys[(size_t)k * TC_M + tid] = v;
                s = fma(v, v, s);
            }
            yy_s[tid] = s;
        }
        __syncthreads();

        const int64_t gw = tile * 4 + warp;  // global mask word of this warp's 32 bases
        if (gw >= P.W) continue;             // warp-uniform (tail tile)
        const int64_t rem = P.N - gw * 32;
        const uint32_t valid_bits = rem >= 32 ? 0xffffffffu : ((1u << rem) - 1u);
        const uint32_t* crow = P.cand + (size_t)gw * P.Bpad;
        const double* ycol = ys + warp * 32;
        const double* yyw = yy_s + warp * 32;

        for (int64_t b0 = 0; b0 < P.B; b0 += 32) {
            const int64_t b = b0 + lane;
            const uint32_t cw = (b < P.B) ? (crow[b] & valid_bits) : 0u;
            // work list, sorted by (permutation, base)
            const int c = __popc(cw);
            int inc = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += t;
            }
            const int nwork = __shfl_sync(0xffffffffu, inc, 31);
            if (nwork == 0) {  // nothing to refine for these 32 permutations
                if (b < P.B) {
                    P.mask[b * P.W + gw] = 0u;
                    P.off[b * P.W + gw] = 0;
                }
                continue;
            }
            ew_w[lane] = 0u;
            {
                uint32_t t = cw;
                int o = inc - c;
                while (t) {
                    const int bit = __ffs(t) - 1;
                    t &= t - 1;
                    list_w[o++] = (uint16_t)((lane << 5) | bit);
                }
            }
            // permutation rows of this block as bytes (n <= 255), coalesced over the [32][n] ints
            {
                const int64_t nb = min((int64_t)32, P.B - b0);
                const int32_t* src = P.idx + b0 * n;
                for (int e = lane; e < (int)nb * n; e += 32) ix_w[e] = (unsigned char)src[e];
            }
            if (nwork > chunk_left) {  // warp-uniform: grab a fresh chunk (the remainder is left unused)
                unsigned long long g = 0;
                if (lane == 0) g = atomicAdd(P.cursor, (unsigned long long)RF_CHUNK);
                chunk_ptr = __shfl_sync(0xffffffffu, g, 0);
                chunk_left = RF_CHUNK;
            }
            __syncwarp();
            unsigned long long run = chunk_ptr;
            for (int e0 = 0; e0 < nwork; e0 += 32) {
                const int e = e0 + lane;
                bool pass = false;
                double f = 0.0;
                int q = 0, bl = 0;
                if (e < nwork) {
                    const int ent = list_w[e];
                    q = ent >> 5;
                    bl = ent & 31;
                    const unsigned char* ix = ix_w + q * n;
                    double s0 = 0.0, sd = 0.0;
                    // S_j in k order, s0 / sd in j order: identical chains to the dense kernel
                    if (PT > 0) {
                        constexpr int QS = (PT + 1) & ~1;
                        double S[PT > 0 ? PT : 1];
#pragma unroll
                        for (int j = 0; j < PT; ++j) S[j] = 0.0;
#pragma unroll 4
                        for (int k = 0; k < n; ++k) {
                            const double yv = ycol[(size_t)k * TC_M + bl];
                            const double2* qrow = reinterpret_cast<const double2*>(Qs + (size_t)ix[k] * QS);
#pragma unroll
                            for (int j2 = 0; j2 < QS / 2; ++j2) {
                                const double2 qq = qrow[j2];
                                S[2 * j2] = fma(yv, qq.x, S[2 * j2]);
                                if (2 * j2 + 1 < PT) S[2 * j2 + 1] = fma(yv, qq.y, S[2 * j2 + 1]);
                            }
                        }
#pragma unroll
                        for (int j = 0; j < PT; ++j) {
                            if (j < P.p0) s0 = fma(S[j], S[j], s0);
                            else sd = fma(S[j], S[j], sd);
                        }
                    } else {
                        for (int j = 0; j < p; ++j) {
                            double S = 0.0;
                            for (int k = 0; k < n; ++k)
                                S = fma(ycol[(size_t)k * TC_M + bl], Qs[(size_t)ix[k] * qs + j], S);
                            if (j < P.p0) s0 = fma(S, S, s0);
                            else sd = fma(S, S, sd);
                        }
                    }
                    double rss1 = __dsub_rn(__dsub_rn(yyw[bl], s0), sd);
                    if (rss1 < 0.0) rss1 = 0.0;
                    const double num = num_is_sd ? sd : __ddiv_rn(sd, P.dfnum);
                    const double den = __dadd_rn(P.adjust_f, __ddiv_rn(rss1, P.dfden));
                    f = __ddiv_rn(num, den);
                    pass = f >= P.U;
                }
                const unsigned bal = __ballot_sync(0xffffffffu, pass);
                if (pass) {
                    const unsigned long long dst = run + __popc(bal & ((1u << lane) - 1u));
                    if (dst < P.capacity) P.vals[dst] = f;
                    atomicOr(&ew_w[q], 1u << bl);
                }
                run += __popc(bal);
            }
            __syncwarp();
            // lane = permutation again: exact word and the offset of its first value
            const uint32_t ew = ew_w[lane];
            const int c2 = __popc(ew);
            int inc2 = c2;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, inc2, o);
                if (lane >= o) inc2 += t;
            }
            if (b < P.B) {
                P.mask[b * P.W + gw] = ew;
                P.off[b * P.W + gw] = (int64_t)(chunk_ptr + (unsigned long long)(inc2 - c2));
            }
            chunk_left -= (int)(run - chunk_ptr);
            chunk_ptr = run;
            __syncwarp();
        }
    }
}

// ==========================================================================================
// Fused null scan: screen + refinement in ONE persistent kernel (the default K2 null path).
//
// A CTA walks 128-base tiles.  Per tile the 4 epilogue warps build (i) the centred FP64 tile in
// shared memory (exact contract, read from HBM once) and (ii) its bf16 [hi | lo] split as the UMMA
// A operand; warp 4 streams the pre-built B tiles (bulk copy + mbarrier, 2 stages) and issues the
// MMAs  hi*qh + hi*ql + lo*qh  into a double-buffered TMEM accumulator.  Each epilogue warp then
// owns 32 bases: it reads its TMEM lanes, ballots the candidate test into one word per
// permutation, releases the accumulator, and immediately refines the candidates of that
// permutation chunk with the FP64 contract (same code path as fstat_refine_kernel), writing the
// exact mask word, the value offset and the compacted values.  Nothing but the final results
// touches HBM.

struct FusedGeom {
    int n, npad, nk16;  // samples, padded to 16, K steps of 16 per operand term
    int KB;             // 64-element K blocks of the stored [hi | lo] rows (2*npad elements)
    int p, p0, j0, pe, pst;  // model columns, first column used by the screen, columns screened, TMEM stride
    int ps, nmma;       // permutations per chunk, MMA N = ps * pst
    int tmem_cols;      // TMEM allocation: nt accumulator stages, rounded up to a power of two
    int nt;             // accumulator stages (2; 1 trades the intra-CTA overlap for a second CTA per SM)
    int qs, ixs;        // Qs row stride (doubles), bytes per row of the byte permutation table
    int nb;             // B tile stages: 1 when the tile is small (occupancy first), else 2
    int ew;             // epilogue warps per TMEM lane quarter (1, 2 or 4)
    int ys_smem;        // 1: centred FP64 tile resident in shared memory; 0: streamed (candidates re-read Y)
    int64_t B, chunks;
    int64_t tiles;
    uint32_t a_bytes, b_bytes;
    uint32_t off_b, off_ys, off_qs, off_yy, off_thr, off_bar, off_warp, warp_bytes;
    size_t smem;
};

// B operand of the fused kernel: per chunk a [nmma x KB*64] bf16 tile, rows r = q*pst + jj (column
// j0 + jj of permutation q, pre-scaled by sqrt(w_j)), K layout [qh (npad) | ql (npad) | 0].
__global__ void __launch_bounds__(128) build_qb2_kernel(const double* __restrict__ q /*[n][p]*/,
                                                        const int32_t* __restrict__ idx /*[B][n]*/, FusedGeom G,
                                                        double sw0, double swd, unsigned char* __restrict__ out,
                                                        unsigned char* __restrict__ idx8) {
    const int64_t chunks_per_row = (int64_t)G.KB * 8;
    const int64_t total = G.chunks * G.nmma * chunks_per_row;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = t / (G.nmma * chunks_per_row);
        const int64_t rem = t - c * G.nmma * chunks_per_row;
        const int row = (int)(rem / chunks_per_row);
        const int kc = (int)(rem - (int64_t)row * chunks_per_row);
        const int qi = row / G.pst, jj = row - qi * G.pst;
        const int j = G.j0 + jj;
        const int64_t b = c * G.ps + qi;
        __nv_bfloat16 vals[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int col = kc * 8 + e;
            const int term = col / G.npad, k = col - term * G.npad;
            float f = 0.f;
            if (b < G.B && jj < G.pe && term < 2 && k < G.n) {
                const double v = q[(size_t)idx[b * G.n + k] * G.p + j] * (j < G.p0 ? sw0 : swd);
                const __nv_bfloat16 h = __double2bfloat16(v);
                f = term == 0 ? __bfloat162float(h) : __bfloat162float(__double2bfloat16(v - (double)__bfloat162float(h)));
            }
            vals[e] = __float2bfloat16(f);
        }
        const int kb = (kc * 8) / 64, colb = (kc * 8) % 64;
        unsigned char* dst = out + ((size_t)(c * G.KB + kb) * G.nmma) * 128 + sw128_offset(row, colb);
        *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(vals);
    }
    // permutation rows as bytes, padded to ixs per row (rows beyond B are zero)
    const int64_t rows = G.chunks * G.ps + 32;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < rows * G.ixs; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = t / G.ixs;
        const int k = (int)(t - b * G.ixs);
        idx8[t] = (b < G.B && k < G.n) ? (unsigned char)idx[b * G.n + k] : 0;
    }
}

struct FusedParams {
    const double* Y;
    int64_t ld, N;
    FusedGeom G;
    const unsigned char* qb;    // [chunks][KB][nmma][128 B]
    const unsigned char* idx8;  // [chunks*ps + 32][ixs]
    const double* q;            // [n][p] row-major FP64
    int center;
    double U, adjust_f, dfnum, dfden;
    float kd;    // kappa * delta
    float e0;    // sqrt(c) * 2 * n^1.5 * 2^-52 when the constant column is dropped from the screen, else 0
    uint32_t idesc;
    uint32_t* mask;  // [B][W]
    int64_t W;
    int64_t* off;    // [B][W]
    double* vals;
    unsigned long long capacity;
    unsigned long long* cursor;
    unsigned long long* prof;  // [16] cycle counters (dbg & 64)
    int dbg;  // developer experiments (timing only, results wrong): 1 = reuse the first B tiles, 2 = no MMAs
};

// permutations per epilogue pass for a column stride of `pcols` (host and device agree on this)
__host__ __device__ constexpr int fz_pass_perms(int pcols) { return pcols <= 4 ? 16 : (pcols <= 8 ? 8 : 2); }

// PT: number of model columns (2..8) or 0 = generic; J0: 1 when the constant column is not screened
// Thread layout: EW epilogue warps per 32-lane TMEM quarter (warp w: quarter w & 3, share h = w >> 2
// of the chunk's permutations), then one producer / MMA warp (index 4 * EW).  With one CTA per SM
// (large sample counts: the FP64 tile fills shared memory) a single epilogue warp per scheduler
// cannot hide its own latencies and the tensor pipe starves; EW = 4 gives every scheduler four.
// The siblings of a quarter work on the same 32 bases (same centred tile), different permutations.
constexpr int FZ_MAX_THREADS = 4 * 128 + 32;
template <int PT, int J0>
__global__ void __launch_bounds__(FZ_MAX_THREADS) fstat_null_kernel(const FusedParams P) {
    constexpr int PCOLS = PT > 0 ? PT - J0 : 16;
    // TMEM column stride per permutation = the number of screened columns: no padded columns in the
    // B tile, the MMA or the accumulator (5 of 8 columns used to be real for p = 6)
    constexpr int PST = PCOLS;
    constexpr int PP = fz_pass_perms(PCOLS);  // permutations per epilogue pass
    constexpr int NC = PP * PCOLS;            // TMEM columns per pass (multiple of 8, <= 64)
    extern __shared__ __align__(1024) unsigned char smem_fz_raw[];
    unsigned char* sm = smem_fz_raw + ((1024u - (smem_u32(smem_fz_raw) & 1023u)) & 1023u);
    const FusedGeom& G = P.G;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n = G.n, p = G.p, qs = G.qs;
    const int ew = G.ew, mma_warp = 4 * G.ew;
    const int quarter = warp & 3, hsub = warp >> 2;  // epilogue warps only
    const int hp = G.ps / ew;                        // permutations of a chunk handled by one warp
    const int row = quarter * 32 + lane;             // base row (TMEM lane) of this thread
    unsigned char* As = sm;
    unsigned char* Bs = sm + G.off_b;
    double* ys = reinterpret_cast<double*>(sm + G.off_ys);  // [n][128] centred
    double* Qs = reinterpret_cast<double*>(sm + G.off_qs);  // [n][qs]
    double* yy_s = reinterpret_cast<double*>(sm + G.off_yy);  // [128]
    double* m_s = yy_s + TC_M;                                // [128] row means (streamed-Y variant)
    const bool ys_smem = G.ys_smem != 0;
    float* thr_s = reinterpret_cast<float*>(sm + G.off_thr);  // [128] candidate thresholds
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + G.off_bar);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);
    uint64_t* mma_a = bars + 14;   // [3 * nk16] A descriptors of a chunk's MMA sequence
    uint64_t* mma_b = mma_a + 48;  // [3 * nk16] B descriptors relative to stage 0
    const uint32_t full_b = smem_u32(bars + 0), empty_b = smem_u32(bars + 4), a_ready = smem_u32(bars + 12);
    const uint32_t tmem_full = smem_u32(bars + 8), tmem_empty = smem_u32(bars + 10);
    const int NB = G.nb;  // B stages (1 .. 4)
    // per-warp scratch (epilogue warps only)
    unsigned char* wbase = sm + G.off_warp + (size_t)(warp < mma_warp ? warp : 0) * G.warp_bytes;
    uint16_t* list_w = reinterpret_cast<uint16_t*>(wbase);              // [hp * 32] (perm << 5) | base
    uint32_t* ew_w = reinterpret_cast<uint32_t*>(list_w + hp * 32);     // [32] exact words
    unsigned char* ix_w = reinterpret_cast<unsigned char*>(ew_w + 32);  // [hp][ixs] permutation rows

    if (warp == mma_warp) {
        if (lane == 0) {
            mbar_init(a_ready, 4);
            for (int s = 0; s < NB; ++s) {
                mbar_init(full_b + 8 * s, 1);
                mbar_init(empty_b + 8 * s, 1);
            }
            for (int s = 0; s < 2; ++s) {
                mbar_init(tmem_full + 8 * s, 1);
                mbar_init(tmem_empty + 8 * s, (uint32_t)(4 * ew));
            }
            fence_barrier_init();
            // MMA sequence of a chunk: for every 16-wide K step  hi*qh, hi*ql, lo*qh
            const uint32_t a_addr0 = smem_u32(As), b_addr0 = smem_u32(Bs);
            for (int t = 0; t < G.nk16; ++t)
                for (int term = 0; term < 3; ++term) {
                    const int acol = (term == 2 ? G.npad : 0) + 16 * t;
                    const int bcol = (term == 1 ? G.npad : 0) + 16 * t;
                    mma_a[3 * t + term] =
                        umma_desc_sw128(a_addr0 + (uint32_t)(acol >> 6) * TC_M * 128u + (uint32_t)(acol & 63) * 2u);
                    mma_b[3 * t + term] =
                        umma_desc_sw128(b_addr0 + (uint32_t)(bcol >> 6) * G.nmma * 128u + (uint32_t)(bcol & 63) * 2u);
                }
        }
        __syncwarp();
        tmem_alloc(smem_u32(tmem_slot), (uint32_t)G.tmem_cols);
    } else {
        for (int e = tid; e < n * qs; e += 128 * ew) {
            const int k = e / qs, j = e - k * qs;
            Qs[e] = j < p ? P.q[(size_t)k * p + j] : 0.0;
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    unsigned long long chunk_ptr = 0;  // warp-uniform: next free value slot of the warp's chunk
    int chunk_left = 0;
    const bool num_is_sd = P.dfnum == 1.0;
    // developer instrumentation (dbg & 64): cycles per phase, summed over the kernel
    const bool prof = (P.dbg & 64) != 0;
    long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tprev = clock64();
    auto lap = [&](int slot) {
        if (prof) {
            const long long t = clock64();
            pc[slot] += t - tprev;
            tprev = t;
        }
    };
    int64_t it = 0;  // chunks issued / consumed so far (TMEM stage = it & 1, phase = (it >> 1) & 1)
    uint32_t tphase = 0;  // parity of the A-tile barrier (one phase per tile)
    const uint32_t lt_mask = (1u << lane) - 1u;
    // producer state (MMA lane only): incremental stage indices and barrier parities
    struct {
        int sb = 0, s = 0;               // B stage / accumulator stage of the next chunk
        uint32_t full_par = 0;           // bit sb: parity the next wait on full_b[sb] expects
        uint32_t empty_par = 0;          // bit sb: parity the next refill wait on empty_b[sb] expects
        uint32_t acc_par = 0;            // bit s: parity of the accumulator hand-off
        int load_stage = 0;              // stage the next refill goes to
        int load_stage_fill = 0;         // prologue counter
        int load_chunk = 0;              // permutation chunk the next load fetches (B does not depend on the tile)
        int64_t loads_left = 0;          // bulk copies this CTA still has to issue
        bool started = false;
    } pl;
    {
        const int64_t my_tiles = blockIdx.x < G.tiles ? (G.tiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;
        pl.loads_left = my_tiles * G.chunks;
    }
    for (int64_t tile = blockIdx.x; tile < G.tiles; tile += gridDim.x) {
        const int64_t base0 = tile * TC_M;
        float thr = __int_as_float(0x7f800000);  // +inf: rows beyond N never pass
        // siblings of a quarter share its centred tile: they meet before it is overwritten ...
        if (ew > 1 && warp < mma_warp) asm volatile("bar.sync %0, %1;" ::"r"(1 + quarter), "r"(32 * ew) : "memory");
        if (warp < 4) {
            // pull this CTA's NEXT tile towards L2 while the current one is processed: its staging then
            // pays an L2 hit instead of a DRAM round trip per load batch (one 128-byte line per thread)
            if (n >= 32) {  // measured: pays for wide tiles (C3s -4 %), costs issue slots for n = 12
                const int64_t nbase = (tile + gridDim.x) * TC_M;
                for (int i = tid; i < n * 8; i += TC_M) {
                    const int k = i >> 3, line = i & 7;
                    const int64_t b = nbase + line * 16;
                    if (b < P.N) asm volatile("prefetch.global.L2 [%0];" ::"l"(P.Y + (size_t)k * P.ld + b));
                }
            }
            // ---- stage the tile: FP64 centred copy (exact contract) + bf16 hi/lo A operand
            const int64_t base = base0 + tid;
            const bool valid = base < P.N && !(P.dbg & 32);
            double acc = 0.0;
            if (valid) {
                constexpr int LB = 12;  // independent loads in flight, then the ordered sum
                for (int k0 = 0; k0 < n; k0 += LB) {
                    double v[LB];
#pragma unroll
                    for (int e = 0; e < LB; ++e) v[e] = (k0 + e < n) ? P.Y[(size_t)(k0 + e) * P.ld + base] : 0.0;
#pragma unroll
                    for (int e = 0; e < LB; ++e)
                        if (k0 + e < n) {
                            if (ys_smem) ys[(size_t)(k0 + e) * TC_M + tid] = v[e];
                            acc = (k0 + e == 0) ? v[e] : __dadd_rn(acc, v[e]);
                        }
                }
            } else if (ys_smem) {
                for (int k = 0; k < n; ++k) ys[(size_t)k * TC_M + tid] = 0.0;
            }
            const double m = P.center ? __ddiv_rn(acc, (double)n) : 0.0;
            double yy = 0.0;
            auto put = [&](int col, const uint4& val) {  // col: multiple of 8
                *reinterpret_cast<uint4*>(As + (size_t)(col >> 6) * TC_M * 128 + sw128_offset(tid, col & 63)) = val;
            };
            for (int k0 = 0; k0 < G.npad; k0 += 8) {
                __nv_bfloat162 hi[4], lo[4];
                double raw[8];
                if (!ys_smem) {  // streamed-Y variant: the second pass re-reads the (L2-hot) values
#pragma unroll
                    for (int e = 0; e < 8; ++e)
                        raw[e] = (valid && k0 + e < n) ? P.Y[(size_t)(k0 + e) * P.ld + base] : 0.0;
                }
#pragma unroll
                for (int e = 0; e < 8; e += 2) {
                    float f[2];
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const int k = k0 + e + u;
                        double v = 0.0;
                        if (k < n) {
                            if (ys_smem) {
                                v = __dsub_rn(ys[(size_t)k * TC_M + tid], m);
                                ys[(size_t)k * TC_M + tid] = v;
                            } else {
                                v = valid ? __dsub_rn(raw[e + u], m) : 0.0;
                            }
                            yy = fma(v, v, yy);
                        }
                        f[u] = (float)v;
                    }
                    const __nv_bfloat162 h = __floats2bfloat162_rn(f[0], f[1]);
                    const float2 hf = __bfloat1622float2(h);
                    hi[e >> 1] = h;
                    lo[e >> 1] = __floats2bfloat162_rn(f[0] - hf.x, f[1] - hf.y);  // exact differences
                }
                put(k0, *reinterpret_cast<const uint4*>(hi));
                put(G.npad + k0, *reinterpret_cast<const uint4*>(lo));
            }
            const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
            for (int col = 2 * G.npad; col < G.KB * 64; col += 8) put(col, zero4);
            yy_s[tid] = yy;
            if (!ys_smem) m_s[tid] = m;
            if (valid) {
                // candidate iff  L~ >= (sqrt(T) - kappa*delta*|yc| - e0*(|m| + |yc|))^2, T = U*(adjustF + yy/dfden);
                // evaluated in FP32 with the rounding pushed to the safe side
                const float yn = __fsqrt_ru((float)yy * 1.0000002f);
                const float Tf = (float)(P.U * (P.adjust_f + yy / P.dfden));
                const float r = __fsqrt_rd(Tf) * 0.999999f - P.kd * yn - P.e0 * (fabsf((float)m) * 1.0000002f + yn);
                float t2 = r > 0.f ? r * r * 0.99999f : 0.f;  // 1e-5: fp32 epilogue rounding
                thr = (t2 == t2) ? t2 : __int_as_float(0x7fc00000);  // NaN data -> never a candidate
            }
            if (ew > 1) thr_s[tid] = thr;
        }
        if (warp < 4) {
            // A rows of this warp are written: publish them to the async proxy and tell the MMA warp.
            // Epilogue warps never wait for each other -- they only meet through tmem_full.
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_ready);
            lap(0);  // tile staging
        }
        // ... and after it has been rebuilt
        if (ew > 1 && warp < mma_warp) {
            asm volatile("bar.sync %0, %1;" ::"r"(1 + quarter), "r"(32 * ew) : "memory");
            thr = thr_s[row];
        }

        if (warp == mma_warp) {
            // ===== producer + MMA issuer (one lane).  This single thread is on the critical path of every
            // chunk, so its loop carries no divisions and no descriptor arithmetic: stage indices and
            // barrier parities are incremental, the MMA descriptors come from a table built once.
            // B tiles: the first NB are requested up front; afterwards the stage of the PREVIOUS chunk is
            // refilled right after the MMAs of the current one have been issued, so waiting for the
            // previous chunk's MMAs to retire overlaps the execution of the current ones. =====
            if (lane == 0) {
                const uint32_t b_addr = smem_u32(Bs);
                const unsigned char* src = P.qb;
                if (it == 0) {  // first tile of this CTA: prologue loads
                    while (pl.loads_left > 0 && pl.load_stage_fill < NB) {
                        const int sb1 = pl.load_stage_fill;
                        mbar_expect_tx(full_b + 8 * sb1, G.b_bytes);
                        bulk_g2s(b_addr + sb1 * G.b_bytes, src + (size_t)pl.load_chunk * G.b_bytes, G.b_bytes,
                                 full_b + 8 * sb1);
                        pl.load_chunk = pl.load_chunk + 1 == (int)G.chunks ? 0 : pl.load_chunk + 1;
                        ++pl.load_stage_fill;
                        --pl.loads_left;
                    }
                }
                const int nmm = (P.dbg & 2) ? 3 : 3 * G.nk16;
                for (int64_t c = 0; c < G.chunks; ++c) {
                    if (c == 0) {  // the B tiles are in flight while the A tile is still being built
                        lap(7);
                        mbar_wait(a_ready, tphase);
                        tc_fence_after();
                        lap(1);  // MMA lane: wait for the A tile
                    }
                    const int sb = pl.sb, s = pl.s;
                    mbar_wait_sel(full_b + 8 * sb, (pl.full_par >> sb) & 1u, P.dbg & 16);
                    pl.full_par ^= 1u << sb;
                    lap(2);  // MMA lane: wait for the B tile
                    mbar_wait_sel(tmem_empty + 8 * s, ((pl.acc_par >> s) & 1u) ^ 1u, P.dbg & 16);
                    pl.acc_par ^= 1u << s;
                    tc_fence_after();
                    lap(3);  // MMA lane: wait for a free accumulator
                    const uint32_t d_addr = tmem_base + (uint32_t)(s * G.nmma);
                    const uint64_t bstage = (uint64_t)(((uint32_t)sb * G.b_bytes) >> 4);  // start-address field offset
#pragma unroll 3
                    for (int i = 0; i < nmm; ++i) umma_bf16(d_addr, mma_a[i], mma_b[i] + bstage, P.idesc, i ? 1u : 0u);
                    umma_commit(empty_b + 8 * sb);   // B stage reusable once these MMAs have read it
                    umma_commit(tmem_full + 8 * s);  // accumulator ready for the epilogue
                    lap(4);  // MMA lane: issue
                    // refill the stage of the previous chunk (of this one when there is a single stage)
                    if (pl.loads_left > 0 && (NB == 1 || pl.started)) {
                        const int sb1 = pl.load_stage;
                        mbar_wait(empty_b + 8 * sb1, (pl.empty_par >> sb1) & 1u);  // its MMAs have retired
                        pl.empty_par ^= 1u << sb1;
                        if (P.dbg & 1) {
                            mbar_arrive(full_b + 8 * sb1);  // experiment: keep the tile already in the stage
                        } else {
                            mbar_expect_tx(full_b + 8 * sb1, G.b_bytes);
                            bulk_g2s(b_addr + sb1 * G.b_bytes, src + (size_t)pl.load_chunk * G.b_bytes, G.b_bytes,
                                     full_b + 8 * sb1);
                        }
                        pl.load_chunk = pl.load_chunk + 1 == (int)G.chunks ? 0 : pl.load_chunk + 1;
                        pl.load_stage = sb1 + 1 == NB ? 0 : sb1 + 1;
                        --pl.loads_left;
                    }
                    pl.started = true;
                    pl.sb = sb + 1 == NB ? 0 : sb + 1;
                    pl.s = (G.nt > 1) ? (s ^ 1) : 0;
                    lap(5);  // MMA lane: refill (waits for the previous chunk's MMAs to retire)
                }
            }
            __syncwarp();
        } else {
            // ===== epilogue + refinement: warp owns 32 bases (TMEM lanes 32*warp .. +31) =====
            const int64_t gw = tile * 4 + quarter;  // global mask word of this warp's 32 bases
            const bool active = gw < P.W;
            const double* ycol = ys + quarter * 32;  // rows beyond N carry thr = +inf: never candidates
            const double* yyw = yy_s + quarter * 32;
            const double* mw = m_s + quarter * 32;
            const double* yglob = P.Y + base0 + quarter * 32;  // streamed-Y variant: candidates re-read Y
            for (int64_t c = 0; c < G.chunks; ++c) {
                const int64_t i0 = it + c;
                const int s = G.nt > 1 ? (int)(i0 & 1) : 0;
                const uint32_t ph = G.nt > 1 ? (uint32_t)((i0 >> 1) & 1) : (uint32_t)(i0 & 1);
                const int64_t b0 = c * G.ps + hsub * hp;  // first permutation of this warp's share of the chunk
                const int nq = (int)max((int64_t)0, min((int64_t)hp, P.G.B - b0));  // valid permutations of the share
                lap(3);  // epilogue: refinement + bookkeeping of the previous chunk
                mbar_wait_sel(tmem_full + 8 * s, ph, P.dbg & 16);
                tc_fence_after();
                lap(1);  // epilogue: wait for the accumulator
                const uint32_t taddr =
                    tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(s * G.nmma + hsub * hp * PST);
                // candidate test + work list in one pass: the ballot word of a permutation is warp-uniform,
                // so its candidates are appended (in base order) right away -> list sorted by (perm, base)
                // All statistics of a 32-column load are formed first (independent FMA chains), then all
                // ballots, and only then the (rare for stringent cutoffs) non-empty words are appended: no
                // vote -> branch dependency between consecutive permutations.
                int nwork = 0;
                for (int q0 = 0; q0 < ((P.dbg & 4) ? 0 : nq); q0 += PP) {
                    uint32_t vr[NC];
                    tmem_ld_cols<NC>(taddr + (uint32_t)(q0 * PST), vr);
                    const bool full = q0 + PP <= nq;  // warp-uniform: no padded permutation in this pass
                    unsigned words[PP];
                    unsigned any = 0u;
#pragma unroll
                    for (int qq = 0; qq < PP; ++qq) {
                        const float v0 = __uint_as_float(vr[qq * PST]);
                        float a = v0 * v0;
#pragma unroll
                        for (int j = 1; j < PCOLS; ++j) {
                            const float vj = __uint_as_float(vr[qq * PST + j]);
                            a = fmaf(vj, vj, a);
                        }
                        words[qq] = __ballot_sync(0xffffffffu, a >= thr);
                    }
                    if (!full) {
#pragma unroll
                        for (int qq = 0; qq < PP; ++qq)
                            if (q0 + qq >= nq) words[qq] = 0u;
                    }
#pragma unroll
                    for (int qq = 0; qq < PP; ++qq) any |= words[qq];
                    if (any) {
                        const uint16_t ent0 = (uint16_t)((q0 << 5) | lane);
#pragma unroll
                        for (int qq = 0; qq < PP; ++qq) {
                            const unsigned word = words[qq];
                            if (word) {
                                if ((word >> lane) & 1u) list_w[nwork + __popc(word & lt_mask)] = (uint16_t)(ent0 + (qq << 5));
                                nwork += __popc(word);
                            }
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tmem_empty + 8 * s);  // accumulator free: the next MMAs overlap the refinement
                lap(2);  // epilogue: TMEM loads + candidate scan
                const int64_t b = b0 + lane;
                if (nwork == 0) {
                    if (lane < nq && active && !(P.dbg & 8)) {
                        P.mask[b * P.W + gw] = 0u;
                        P.off[b * P.W + gw] = 0;
                    }
                    __syncwarp();
                    continue;
                }
                ew_w[lane] = 0u;
                {  // permutation rows of this chunk as bytes: nq rows of ixs bytes, 16-byte copies
                    const uint4* src = reinterpret_cast<const uint4*>(P.idx8 + (size_t)b0 * G.ixs);
                    const int nv = nq * (G.ixs >> 4);
                    for (int e = lane; e < nv; e += 32) reinterpret_cast<uint4*>(ix_w)[e] = src[e];
                }
                if (nwork > chunk_left) {  // warp-uniform: grab a fresh chunk (the remainder is left unused)
                    unsigned long long g = 0;
                    if (lane == 0) g = atomicAdd(P.cursor, (unsigned long long)RF_CHUNK);
                    chunk_ptr = __shfl_sync(0xffffffffu, g, 0);
                    chunk_left = RF_CHUNK;
                }
                __syncwarp();
                unsigned long long run = chunk_ptr;
                for (int e0 = 0; e0 < nwork; e0 += 32) {
                    const int e = e0 + lane;
                    bool pass = false;
                    double f = 0.0;
                    int q = 0, bl = 0;
                    if (e < nwork) {
                        const int ent = list_w[e];
                        q = ent >> 5;
                        bl = ent & 31;
                        const unsigned char* ix = ix_w + q * G.ixs;
                        double s0 = 0.0, sd = 0.0;
                        // S_j in k order, s0 / sd in j order: identical chains to the dense kernel
                        if (PT > 0) {
                            constexpr int QS = (PT + 1) & ~1;
                            double S[PT > 0 ? PT : 1];
#pragma unroll
                            for (int j = 0; j < PT; ++j) S[j] = 0.0;
                            const double mrow = ys_smem ? 0.0 : mw[bl];
#pragma unroll 4
                            for (int k = 0; k < n; ++k) {
                                const double yv = ys_smem ? ycol[(size_t)k * TC_M + bl]
                                                          : __dsub_rn(yglob[(size_t)k * P.ld + bl], mrow);
                                const double2* qrow = reinterpret_cast<const double2*>(Qs + (size_t)ix[k] * QS);
#pragma unroll
                                for (int j2 = 0; j2 < QS / 2; ++j2) {
                                    const double2 qq = qrow[j2];
                                    S[2 * j2] = fma(yv, qq.x, S[2 * j2]);
                                    if (2 * j2 + 1 < PT) S[2 * j2 + 1] = fma(yv, qq.y, S[2 * j2 + 1]);
                                }
                            }
#pragma unroll
                            for (int j = 0; j < PT; ++j) {
                                if (j < G.p0) s0 = fma(S[j], S[j], s0);
                                else sd = fma(S[j], S[j], sd);
                            }
                        } else {
                            const double mrow = ys_smem ? 0.0 : mw[bl];
                            for (int j = 0; j < p; ++j) {
                                double S = 0.0;
                                for (int k = 0; k < n; ++k) {
                                    const double yv = ys_smem ? ycol[(size_t)k * TC_M + bl]
                                                              : __dsub_rn(yglob[(size_t)k * P.ld + bl], mrow);
                                    S = fma(yv, Qs[(size_t)ix[k] * qs + j], S);
                                }
                                if (j < G.p0) s0 = fma(S, S, s0);
                                else sd = fma(S, S, sd);
                            }
                        }
                        double rss1 = __dsub_rn(__dsub_rn(yyw[bl], s0), sd);
                        if (rss1 < 0.0) rss1 = 0.0;
                        const double num = num_is_sd ? sd : __ddiv_rn(sd, P.dfnum);
                        const double den = __dadd_rn(P.adjust_f, __ddiv_rn(rss1, P.dfden));
                        f = __ddiv_rn(num, den);
                        pass = f >= P.U;
                    }
                    const unsigned bal = __ballot_sync(0xffffffffu, pass);
                    if (pass) {
                        const unsigned long long dst = run + __popc(bal & ((1u << lane) - 1u));
                        if (dst < P.capacity) P.vals[dst] = f;
                        atomicOr(&ew_w[q], 1u << bl);
                    }
                    run += __popc(bal);
                }
                __syncwarp();
                // lane = permutation again: exact word and the offset of its first value
                const uint32_t ew = ew_w[lane];
                const int c2 = __popc(ew);
                int inc2 = c2;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, inc2, o);
                    if (lane >= o) inc2 += t;
                }
                if (lane < nq) {
                    P.mask[b * P.W + gw] = ew;
                    P.off[b * P.W + gw] = (int64_t)(chunk_ptr + (unsigned long long)(inc2 - c2));
                }
                chunk_left -= (int)(run - chunk_ptr);
                chunk_ptr = run;
                __syncwarp();
            }
        }
        it += G.chunks;
        tphase ^= 1u;
    }
    if (prof && lane == 0 && (warp == 0 || warp == mma_warp) && P.prof) {
        unsigned long long* dst = P.prof + (warp == mma_warp ? 8 : 0);
        for (int i = 0; i < 8; ++i) atomicAdd(&dst[i], (unsigned long long)pc[i]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == mma_warp) tmem_dealloc(tmem_base, (uint32_t)G.tmem_cols);
}

__global__ void popcount_words_kernel(const uint32_t* __restrict__ w, int64_t n, unsigned long long* __restrict__ out) {
    unsigned long long c = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        c += __popc(w[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

// words that differ between two masks of the same shape (debug / validation)
__global__ void mask_diff_kernel(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b, int64_t n,
                                 unsigned long long* __restrict__ out) {
    unsigned long long c = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        c += __popc(a[i] ^ b[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

int mask_diff(dfb_ctx* ctx, const uint32_t* a, const uint32_t* b, int64_t words, int64_t* diff) {
    DevBuf<unsigned long long> cnt;
    DFB_TRY(cnt.alloc(ctx, 1));
    DFB_TRY(cnt.zero());
    mask_diff_kernel<<<grid_for(words, 256, ctx->sm_count, 8), 256, 0, ctx->stream>>>(a, b, words, cnt.p);
    unsigned long long h = 0;
    DFB_TRY(d2h(ctx, &h, cnt.p, sizeof h));
    *diff = (int64_t)h;
    return DFB_OK;
}

// ------------------------------------------------------------------------------------------
// host side

static bool screen_geometry(const dfb_cov* cov, const ModelBasis& mb, int64_t B, size_t smem_limit, ScreenGeom* g) {
    ScreenGeom G;
    memset(&G, 0, sizeof G);
    G.n = cov->n;
    G.npad = (int)round_up(G.n, 16);
    G.Kp = (int)round_up(3 * G.npad, 64);
    G.KB = G.Kp / 64;
    G.p = mb.p;
    G.p0 = mb.p0;
    G.pst = 2;
    while (G.pst < G.p) G.pst *= 2;
    if (G.pst > 16) return false;
    G.B = B;
    G.tiles = ceil_div(cov->N, TC_M);
    for (int nmma = std::min(128, 32 * G.pst); nmma >= 16; nmma /= 2) {
        G.nmma = nmma;
        G.ps = nmma / G.pst;
        G.smem = (size_t)G.KB * TC_M * 128 + 2 * (size_t)G.KB * nmma * 128 + 8 * 8 + 16 + 4 * 32 * 4 + 1024;
        if (G.smem <= smem_limit && G.ps >= 1) {
            G.chunks = ceil_div(B, G.ps);
            G.Bpad = G.chunks * G.ps;
            *g = G;
            return true;
        }
    }
    return false;
}

// conservative bound on |S~ - S| / |yc| for the 3-term bf16 split with fp32 accumulation over
// 3*n products (derivation in DESIGN.md), times a 2x safety factor
static double screen_delta(int n) { return 2.0 * (3.1 * 3.8147e-6 /*2^-18*/ + 3.0 * n * 2.3842e-7 /*2^-22*/); }

bool screen_supported(const dfb_ctx* ctx, const dfb_cov* cov, const ModelBasis& mb, double lo, double hi,
                      double adjust_f) {
    if (!ctx->opt_screen) return false;
    if (!(hi > 0.0) || !(lo < 0.0)) return false;      // one-sided positive cutoff only
    if (cov->n - mb.p <= 0 || !(adjust_f >= 0.0)) return false;
    if (cov->n > 255) return false;  // the refinement keeps permutation rows as bytes
    ScreenGeom g;
    const size_t limit = std::min<size_t>(ctx->smem_optin, 200 * 1024);
    if (!screen_geometry(cov, mb, 1, limit, &g)) return false;
    // refinement tile must fit too
    return refine_smem_bytes(cov->n, mb.p) <= limit;
}

// Screen + refine for one batch of permutations; fills the same PermBatch the dense kernel does.
int fstat_perm_batch_screen(dfb_ctx* ctx, const dfb_cov* cov, const ModelBasis& mb, const double* q_dev /*[n][p]*/,
                            double adjust_f, const int32_t* idx_dev, int64_t B, double hi, size_t capacity_hint,
                            PermBatch* out, int64_t* ncand_out) {
    DFB_TRY(ensure_y(ctx, cov));  // the first-generation kernels read the materialised log2 matrix
    ScreenGeom G;
    const size_t limit = std::min<size_t>(ctx->smem_optin, 200 * 1024);
    if (!screen_geometry(cov, mb, B, limit, &G)) {
        set_error("screen geometry does not fit (internal error)");
        return DFB_EINVAL;
    }
    DevBuf<unsigned char> qb;
    DevBuf<uint32_t> cand;
    const size_t qb_bytes = (size_t)G.chunks * G.KB * G.nmma * 128;
    DFB_TRY(qb.alloc(ctx, qb_bytes));
    DFB_TRY(cand.alloc(ctx, (size_t)G.tiles * 4 * G.Bpad));
    const double df1 = (double)(mb.p - mb.p0), df2 = (double)(cov->n - mb.p);
    const double w_a = 1.0 / df1 + hi / df2, w_c = hi / df2;  // weights of the S_d / S_0 columns
    {
        KTimer t(ctx, DFB_K_FSTAT_TC, 1);
        const int64_t work = G.chunks * G.nmma * (G.Kp / 8);
        build_qb_kernel<<<grid_for(work, 128, ctx->sm_count, 16), 128, 0, ctx->stream>>>(q_dev, idx_dev, G, sqrt(w_c),
                                                                                       sqrt(w_a), qb.p);
    }
    ScreenParams S;
    memset(&S, 0, sizeof S);
    S.Y = cov->y.p;
    S.ld = cov->ld;
    S.N = cov->N;
    S.G = G;
    S.qb = qb.p;
    S.center = mb.centered ? 1 : 0;
    S.U = hi;
    S.adjust_f = adjust_f;
    S.dfnum = (double)(mb.p - mb.p0);
    S.dfden = (double)(cov->n - mb.p);
    S.delta = screen_delta(cov->n);
    S.cand = cand.p;
    S.idesc = umma_idesc_bf16(TC_M, G.nmma);
    auto launch = [&](auto kern) -> int {
        DFB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G.smem));
        KTimer t(ctx, DFB_K_FSTAT_TC, 1);
        kern<<<(unsigned)G.tiles, TC_THREADS, G.smem, ctx->stream>>>(S);
        return DFB_OK;
    };
    switch (G.p) {
        case 2: DFB_TRY(launch(fstat_screen_kernel<2, 2>)); break;
        case 3: DFB_TRY(launch(fstat_screen_kernel<4, 3>)); break;
        case 4: DFB_TRY(launch(fstat_screen_kernel<4, 4>)); break;
        case 5: DFB_TRY(launch(fstat_screen_kernel<8, 5>)); break;
        case 6: DFB_TRY(launch(fstat_screen_kernel<8, 6>)); break;
        case 7: DFB_TRY(launch(fstat_screen_kernel<8, 7>)); break;
        case 8: DFB_TRY(launch(fstat_screen_kernel<8, 8>)); break;
        default: DFB_TRY(launch(fstat_screen_kernel<16, 16>)); break;  // zero-padded columns add 0
    }
    DFB_CUDA(cudaGetLastError());

    out->rows = B;
    out->W = ceil_div(cov->N, 32);
    out->tile_bases = 32;  // the refinement emits one value offset per 32-base word
    out->tiles = out->W;
    DFB_TRY(out->mask.alloc(ctx, (size_t)out->rows * out->W));
    DFB_TRY(out->off.alloc(ctx, (size_t)out->rows * out->tiles));
    DevBuf<unsigned long long> cursor;
    DFB_TRY(cursor.alloc(ctx, 1));
    const size_t dense = (size_t)B * (size_t)cov->N;
    const size_t rf_smem = refine_smem_bytes(cov->n, mb.p);
    // persistent grid: as many CTAs as fit (4 per SM is enough to hide the per-tile staging)
    const int rf_grid = (int)std::min<int64_t>(G.tiles, (int64_t)ctx->sm_count * 4);
    // every warp may leave one partly used chunk behind
    const size_t slack = (size_t)rf_grid * 4 * RF_CHUNK;
    size_t cap = std::min(dense, std::max<size_t>(capacity_hint, 1024)) + slack;
    auto launch_refine = [&](const RefineParams& R) -> int {
        auto go = [&](auto kern) -> int {
            DFB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rf_smem));
            KTimer t(ctx, DFB_K_FSTAT, 1);
            kern<<<rf_grid, RF_THREADS, rf_smem, ctx->stream>>>(R);
            return DFB_OK;
        };
        switch (mb.p) {
            case 2: return go(fstat_refine_kernel<2>);
            case 3: return go(fstat_refine_kernel<3>);
            case 4: return go(fstat_refine_kernel<4>);
            case 5: return go(fstat_refine_kernel<5>);
            case 6: return go(fstat_refine_kernel<6>);
            case 7: return go(fstat_refine_kernel<7>);
            case 8: return go(fstat_refine_kernel<8>);
            default: return go(fstat_refine_kernel<0>);
        }
    };
    for (int attempt = 0; attempt < 2; ++attempt) {
        DFB_TRY(out->vals.alloc(ctx, cap));
        DFB_TRY(cursor.zero());
        RefineParams R;
        memset(&R, 0, sizeof R);
        R.Y = cov->y.p;
        R.ld = cov->ld;
        R.N = cov->N;
        R.n = cov->n;
        R.p = mb.p;
        R.p0 = mb.p0;
        R.q = q_dev;
        R.idx = idx_dev;
        R.B = B;
        R.Bpad = G.Bpad;
        R.center = mb.centered ? 1 : 0;
        R.adjust_f = adjust_f;
        R.dfnum = S.dfnum;
        R.dfden = S.dfden;
        R.U = hi;
        R.cand = cand.p;
        R.mask = out->mask.p;
        R.W = out->W;
        R.off = out->off.p;
        R.tiles128 = G.tiles;
        R.vals = out->vals.p;
        R.capacity = cap;
        R.cursor = cursor.p;
        DFB_TRY(launch_refine(R));
        DFB_CUDA(cudaGetLastError());
        unsigned long long used = 0;
        DFB_TRY(d2h(ctx, &used, cursor.p, sizeof used));
        out->nvals = (int64_t)used;  // allocated value slots (exceedances + unused chunk tails)
        if (used <= cap) {
            if (ncand_out) {
                DevBuf<unsigned long long> cntm;
                DFB_TRY(cntm.alloc(ctx, 1));
                DFB_TRY(cntm.zero());
                popcount_words_kernel<<<grid_for(out->rows * out->W, 256, ctx->sm_count, 8), 256, 0, ctx->stream>>>(
                    out->mask.p, out->rows * out->W, cntm.p);
                unsigned long long hm = 0;
                DFB_TRY(d2h(ctx, &hm, cntm.p, sizeof hm));
                out->nvals = (int64_t)hm;  // debug hook: report the exact exceedance count
                DevBuf<unsigned long long> cnt;
                DFB_TRY(cnt.alloc(ctx, 1));
                DFB_TRY(cnt.zero());
                const int64_t words = G.tiles * 4 * G.Bpad;
                popcount_words_kernel<<<grid_for(words, 256, ctx->sm_count, 8), 256, 0, ctx->stream>>>(cand.p, words,
                                                                                                     cnt.p);
                unsigned long long h = 0;
                DFB_TRY(d2h(ctx, &h, cnt.p, sizeof h));
                *ncand_out = (int64_t)h;
            }
            return DFB_OK;
        }
        cap = (size_t)used;
    }
    set_error("exceedance buffer overflow after resize (internal error)");
    return DFB_ECUDA;
}

// ------------------------------------------------------------------------------------------
// fused path, host side

static bool constant_first_column(const ModelBasis& mb) {
    if (!mb.centered || mb.p0 < 1) return false;
    const double c0 = mb.q[0];
    for (int i = 1; i < mb.n; ++i)
        if (fabs(mb.q[(size_t)i * mb.p] - c0) > 1e-14) return false;
    return c0 != 0.0;
}

static bool fused_geometry_ps(const dfb_ctx* ctx, const dfb_cov* cov, const ModelBasis& mb, int64_t B, int force_ps,
                              FusedGeom* g, int force_nt = 0, int force_nb = 0);

// Chunk width.  Small models (<= 4 screened columns: small tiles, several CTAs per SM) use 32
// permutations per chunk.  Larger ones are bound by the tcgen05.mma instruction rate -- an M = 128,
// K = 16 instruction costs ~130 cycles whether N is 80 or 240 (measured) -- so the widest
// chunk whose two accumulator stages fit TMEM (N <= 240) and whose B stages fit shared memory wins.
static bool fused_geometry(const dfb_ctx* ctx, const dfb_cov* cov, const ModelBasis& mb, int64_t B, FusedGeom* g) {
    if (ctx->opt_chunk_perms > 0) return fused_geometry_ps(ctx, cov, mb, B, ctx->opt_chunk_perms, g);
    const int pe = mb.p - ((mb.p <= 8 && constant_first_column(mb)) ? 1 : 0);
    if (mb.p <= 8 && pe >= 5 && ctx->opt_tmem_stages == 0 && ctx->opt_b_stages == 0) {
        // measured best (C3s 5.2 ms): 32-permutation chunks with ONE accumulator and ONE B stage per CTA
        // and two CTAs per SM -- the CTAs overlap each other's tile staging, B loads and hand-offs
        if (32 * pe <= 256 && (32 * pe) % 16 == 0 && fused_geometry_ps(ctx, cov, mb, B, 32, g, 1, 1) &&
            2 * (g->smem + 1024) <= (size_t)227 * 1024)
            return true;
        for (int ps = 48; ps >= 24; ps -= 8)
            if (ps * pe <= 240 && fused_geometry_ps(ctx, cov, mb, B, ps, g)) return true;
    }
    return fused_geometry_ps(ctx, cov, mb, B, 0, g);
}

static bool fused_geometry_ps(const dfb_ctx* ctx, const dfb_cov* cov, const ModelBasis& mb, int64_t B, int force_ps,
                              FusedGeom* g, int force_nt, int force_nb) {
    FusedGeom G;
    memset(&G, 0, sizeof G);
    G.n = cov->n;
    if (G.n > 255) return false;
    G.npad = (int)round_up(G.n, 16);
    G.nk16 = G.npad / 16;
    G.KB = (int)ceil_div(2 * G.npad, 64);
    G.p = mb.p;
    G.p0 = mb.p0;
    G.j0 = (mb.p <= 8 && constant_first_column(mb)) ? 1 : 0;
    G.pe = mb.p - G.j0;
    if (mb.p > 16) return false;
    // TMEM / B-tile column stride per permutation: exactly the screened columns for the templated
    // kernels (p <= 8), 16 (zero padded) for the generic one.  nmma = ps * pst is a multiple of 16.
    G.pst = mb.p > 8 ? 16 : G.pe;
    G.ps = mb.p > 8 ? 8 : (G.pst <= 4 ? 32 : 16);
    if (force_ps > 0) {  // permutations per chunk (multiple of the pass size)
        const int pp = fz_pass_perms(G.pst);
        G.ps = std::max(pp, (force_ps / pp) * pp);
        while (G.ps * G.pst > 256 || (G.ps * G.pst) % 16) G.ps -= pp;
        if (G.ps < pp) return false;
    }
    G.nmma = G.ps * G.pst;
    if (G.nmma < 16 || G.nmma > 256) return false;
    G.tmem_cols = 32;
    G.nt = (force_nt == 1 || ctx->opt_tmem_stages == 1) ? 1 : 2;
    while (G.tmem_cols < G.nt * G.nmma) G.tmem_cols *= 2;
    G.qs = mb.p <= 8 ? ((mb.p + 1) & ~1) : mb.p;
    G.ixs = (int)round_up(G.n, 16);
    G.B = B;
    G.chunks = ceil_div(B, G.ps);
    G.tiles = ceil_div(cov->N, TC_M);
    G.a_bytes = (uint32_t)G.KB * TC_M * 128u;
    G.b_bytes = (uint32_t)G.KB * (uint32_t)G.nmma * 128u;
    // The centred FP64 tile (n x 128 doubles) stays in shared memory when that still leaves room for
    // two CTAs per SM; otherwise it is streamed (the rare candidates re-read Y through L2) so that a
    // second CTA can cover the MMA <-> epilogue hand-off latency and the tile staging of the first.
    {
        const size_t ys_bytes = (size_t)G.n * TC_M * 8;
        const size_t rest = (size_t)G.a_bytes + 2 * (size_t)G.b_bytes + (size_t)G.n * G.qs * 8 + 16 * 1024;
        const bool single = G.tmem_cols > 256;  // the accumulators alone limit the SM to one CTA
        const bool fits = single ? (rest + ys_bytes <= (size_t)225 * 1024) : (2 * (rest + ys_bytes) <= (size_t)225 * 1024);
        G.ys_smem = (ctx->opt_ys_smem >= 0) ? ctx->opt_ys_smem : (fits ? 1 : 0);
        if (!G.ys_smem && rest > (size_t)225 * 1024) return false;
    }
    uint32_t o = G.a_bytes;
    G.off_b = o;
    G.nb = 1;
    if (force_nb >= 1) G.nb = force_nb;
    else if (ctx->opt_b_stages >= 1 && ctx->opt_b_stages <= 4) G.nb = ctx->opt_b_stages;
    else if (G.b_bytes > 16 * 1024) {
        // as many stages as fit (<= 3): a 32 KiB bulk copy out of L2 takes longer than its MMAs
        const size_t fixed = (size_t)G.a_bytes + (G.ys_smem ? (size_t)G.n * TC_M * 8 : 0) + (size_t)G.n * G.qs * 8 + 20 * 1024;
        // streamed tile with small accumulators: leave room for a second CTA
        const size_t budget = (G.ys_smem || G.tmem_cols > 256) ? (size_t)224 * 1024 : (size_t)112 * 1024;
        G.nb = 2;
        while (G.nb < 3 && fixed + (size_t)(G.nb + 1) * G.b_bytes <= budget) ++G.nb;
    }
    o += (uint32_t)G.nb * G.b_bytes;
    G.off_ys = o;
    o += G.ys_smem ? (uint32_t)G.n * TC_M * 8u : 0u;
    G.off_qs = o;
    o += (uint32_t)round_up((int64_t)G.n * G.qs * 8, 16);
    G.off_yy = o;
    o += 2 * TC_M * 8u;  // yy and the row means
    G.off_thr = o;
    o += TC_M * 4u;
    G.off_bar = o;
    o += (14 + 96) * 8 + 16;
    o = (uint32_t)round_up(o, 16);
    G.off_warp = o;
    // epilogue warps per quarter: one when several CTAs share an SM anyway (small tiles); otherwise as
    // many as the 32-column TMEM load granularity allows (a warp's share of a chunk is a multiple of
    // 32 columns), at most 4
    {
        const size_t base_smem = (size_t)o + 4 * (1024 * 2 + 32 * 4 + 32 * G.ixs) + 1024;
        const bool single_cta = 2 * (base_smem + 1024) > (size_t)227 * 1024;
        // measured on B200 (c3s workload): the chunk loop is bound by the depth of the MMA -> epilogue
        // -> MMA hand-off (two TMEM stages), not by epilogue issue, so extra warps do not pay yet; the
        // knob stays for the deeper pipeline ("next")
        (void)single_cta;
        int want = ctx->opt_epi_warps > 0 ? ctx->opt_epi_warps : 1;
        const int pp = fz_pass_perms(G.pst);
        if (ctx->opt_epi_warps <= 0 && G.ps > 32) want = std::min(4, G.ps / 16);  // wide chunks: ~16 permutations per warp
        want = std::max(want, (G.ps + 31) / 32);  // a warp's share of a chunk is at most 32 permutations (lane = permutation)
        G.ew = std::max(1, std::min(std::min(want, 4), G.ps / pp));
        while (G.ew > 1 && (G.ps % G.ew || (G.ps / G.ew) % pp)) --G.ew;
        if (G.ps / G.ew > 32) return false;
    }
    const int hp = G.ps / G.ew;
    G.warp_bytes = (uint32_t)round_up(hp * 32 * 2 + 32 * 4 + hp * G.ixs, 16);
    o += 4 * G.ew * G.warp_bytes;
    G.smem = (size_t)o + 1024;  // slack for the manual 1024-byte alignment
    const size_t limit = std::min<size_t>(ctx->smem_optin, 227 * 1024);
    if (G.smem > limit) return false;
    *g = G;
    return true;
}

bool fused_supported(const dfb_ctx* ctx, const dfb_cov* cov, const ModelBasis& mb, double lo, double hi,
                     double adjust_f) {
    if (!ctx->opt_screen || !ctx->opt_fused) return false;
    if (!(hi > 0.0) || !(lo < 0.0)) return false;  // one-sided positive cutoff only
    if (cov->n - mb.p <= 0 || !(adjust_f >= 0.0)) return false;
    FusedGeom g;
    return fused_geometry(ctx, cov, mb, 1, &g);
}

template <int PT>
static int launch_fused(dfb_ctx* ctx, const FusedParams& F, int grid) {
    auto go = [&](auto kern) -> int {
        DFB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)F.G.smem));
        KTimer t(ctx, DFB_K_FSTAT_TC, 1);
        kern<<<grid, 128 * F.G.ew + 32, F.G.smem, ctx->stream>>>(F);
        return DFB_OK;
    };
    if (PT > 0 && F.G.j0) return go(fstat_null_kernel<PT, PT == 0 ? 0 : 1>);
    return go(fstat_null_kernel<PT, 0>);
}

// Fused screen + refine for one batch of permutations; fills the same PermBatch as the other paths.
int fstat_perm_batch_fused(dfb_ctx* ctx, const dfb_cov* cov, const ModelBasis& mb, const double* q_dev /*[n][p]*/,
                           double adjust_f, const int32_t* idx_dev, int64_t B, double hi, size_t capacity_hint,
                           PermBatch* out, bool defer_check) {
    DFB_TRY(ensure_y(ctx, cov));  // the first-generation kernels read the materialised log2 matrix
    FusedGeom G;
    if (!fused_geometry(ctx, cov, mb, B, &G)) {
        set_error("fused geometry does not fit (internal error)");
        return DFB_EINVAL;
    }
    DevBuf<unsigned char> qb, idx8;
    DFB_TRY(qb.alloc(ctx, (size_t)G.chunks * G.b_bytes));
    DFB_TRY(idx8.alloc(ctx, (size_t)(G.chunks * G.ps + 32) * G.ixs));
    const double df1 = (double)(mb.p - mb.p0), df2 = (double)(cov->n - mb.p);
    const double w_a = 1.0 / df1 + hi / df2, w_c = hi / df2;  // weights of the S_d / S_0 columns
    {
        KTimer t(ctx, DFB_K_FSTAT_TC, 1);
        const int64_t work = G.chunks * G.nmma * (G.KB * 8);
        build_qb2_kernel<<<grid_for(work, 128, ctx->sm_count, 16), 128, 0, ctx->stream>>>(q_dev, idx_dev, G, sqrt(w_c),
                                                                                        sqrt(w_a), qb.p, idx8.p);
    }
    FusedParams F;
    memset(&F, 0, sizeof F);
    F.Y = cov->y.p;
    F.ld = cov->ld;
    F.N = cov->N;
    F.G = G;
    F.qb = qb.p;
    F.idx8 = idx8.p;
    F.q = q_dev;
    F.center = mb.centered ? 1 : 0;
    F.U = hi;
    F.adjust_f = adjust_f;
    F.dfnum = df1;
    F.dfden = df2;
    const double kappa = sqrt(w_a * (double)(mb.p - mb.p0) + w_c * (double)mb.p0);
    F.kd = (float)(kappa * screen_delta(cov->n) * 1.000001);
    // dropped constant column: |S_const| = |sum_k yc[k]| / sqrt(n) <= n^1.5 * 2^-52 * max|y| (rounding of the
    // mean and of the n subtractions); 2x safety.  Its weight in the statistic is c.
    F.e0 = G.j0 ? (float)(sqrt(w_c) * 2.0 * pow((double)cov->n, 1.5) * 2.220446049250313e-16 * 1.000001) : 0.f;
    F.idesc = umma_idesc_bf16(TC_M, G.nmma);
    F.dbg = ctx->opt_dbg;
    DevBuf<unsigned long long> profbuf;
    if (F.dbg & 64) {
        DFB_TRY(profbuf.alloc(ctx, 16));
        DFB_TRY(profbuf.zero());
        F.prof = profbuf.p;
    }

    out->rows = B;
    out->W = ceil_div(cov->N, 32);
    out->tile_bases = 32;
    out->tiles = out->W;
    DFB_TRY(out->mask.alloc(ctx, (size_t)out->rows * out->W));
    DFB_TRY(out->off.alloc(ctx, (size_t)out->rows * out->tiles));
    DevBuf<unsigned long long> cursor;
    DFB_TRY(cursor.alloc(ctx, 1));
    // persistent grid: CTAs per SM limited by shared memory and by TMEM columns (512 per SM)
    const int threads = 128 * G.ew + 32;
    int per_sm = (int)std::min<size_t>((size_t)(227 * 1024) / (G.smem + 1024), (size_t)(512 / G.tmem_cols));
    per_sm = std::min(per_sm, 65536 / (threads * 100));  // registers (the kernels compile to <= 100)
    per_sm = std::max(1, std::min(per_sm, 6));
    // Not one CTA per resident slot for the whole scan: k2_waves (16) times as many, each walking its
    // share of the tiles and retiring.  A fully persistent grid keeps every SM slot until the scan ends,
    // so the observed-region chain on the auxiliary stream (dozens of small dependent kernels) could
    // not start before that; with CTAs retiring every few tiles and the auxiliary stream at the highest
    // priority, its blocks take the freed slots first and the chain finishes inside the scan (C2 step
    // 2.39 -> 2.27 ms; the per-CTA set-up -- TMEM allocation, barriers, Q in shared memory -- is paid 16
    // times per slot and does not show: K2 1.39 ms either way).  Without the priority the extra waves
    // only cost (2.5 ms); with the priority but a persistent grid the scan itself slows down (2.0 ms).
    const int grid = (int)std::min<int64_t>(G.tiles, (int64_t)ctx->sm_count * per_sm * std::max(1, ctx->opt_k2_waves));
    if (getenv("DFB_GAPS"))
        fprintf(stderr, "[dfb] fused geometry: n=%d npad=%d KB=%d pst=%d ps=%d nmma=%d nb=%d ew=%d ys_smem=%d smem=%zu per_sm=%d grid=%d\n",
                G.n, G.npad, G.KB, G.pst, G.ps, G.nmma, G.nb, G.ew, G.ys_smem, G.smem, per_sm, grid);
    const size_t slack = (size_t)grid * 4 * G.ew * RF_CHUNK;  // every warp may leave one partly used chunk behind
    const size_t dense = (size_t)B * (size_t)cov->N;
    size_t cap = std::min(dense, std::max<size_t>(capacity_hint, 1024)) + slack;
    F.mask = out->mask.p;
    F.W = out->W;
    F.off = out->off.p;
    F.cursor = cursor.p;
    for (int attempt = 0; attempt < 2; ++attempt) {
        DFB_TRY(out->vals.alloc(ctx, cap));
        DFB_TRY(cursor.zero());
        F.vals = out->vals.p;
        F.capacity = cap;
        switch (mb.p) {
            case 2: DFB_TRY(launch_fused<2>(ctx, F, grid)); break;
            case 3: DFB_TRY(launch_fused<3>(ctx, F, grid)); break;
            case 4: DFB_TRY(launch_fused<4>(ctx, F, grid)); break;
            case 5: DFB_TRY(launch_fused<5>(ctx, F, grid)); break;
            case 6: DFB_TRY(launch_fused<6>(ctx, F, grid)); break;
            case 7: DFB_TRY(launch_fused<7>(ctx, F, grid)); break;
            case 8: DFB_TRY(launch_fused<8>(ctx, F, grid)); break;
            default: DFB_TRY(launch_fused<0>(ctx, F, grid)); break;
        }
        DFB_CUDA(cudaGetLastError());
        if ((F.dbg & 64) && F.prof) {
            unsigned long long h[16];
            DFB_TRY(d2h(ctx, h, F.prof, sizeof h));
            const double ep = (double)grid, ml = (double)grid;
            fprintf(stderr, "[dfb] fused phases, mean cycles per CTA (epilogue warp 0): staging %.0f  wait-acc %.0f  scan %.0f  refine+rest %.0f\n",
                    h[0] / ep, h[1] / ep, h[2] / ep, h[3] / ep);
            fprintf(stderr, "[dfb] fused phases, mean cycles per CTA (MMA lane): wait-A %.0f  wait-B %.0f  wait-acc %.0f  issue %.0f  refill %.0f  other %.0f\n",
                    h[8 + 1] / ml, h[8 + 2] / ml, h[8 + 3] / ml, h[8 + 4] / ml, h[8 + 5] / ml, h[8 + 7] / ml);
        }
        out->capacity = cap;
        if (defer_check) {
            // the caller synchronises anyway (region counts): it checks *used_host <= capacity then
            out->used_host = d2h_async_pinned(ctx, cursor.p, 1);
            DFB_REQUIRE(out->used_host != nullptr, "pinned read-back of the value cursor failed");
            return DFB_OK;
        }
        unsigned long long used = 0;
        DFB_TRY(d2h(ctx, &used, cursor.p, sizeof used));
        out->nvals = (int64_t)used;  // allocated value slots (exceedances + unused chunk tails)
        if (used <= cap) return DFB_OK;
        cap = (size_t)used;
    }
    set_error("exceedance buffer overflow after resize (internal error)");
    return DFB_ECUDA;
}

}  // namespace dfb
