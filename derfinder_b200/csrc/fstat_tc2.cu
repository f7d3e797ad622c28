// K2 null scan, second generation: one tcgen05 product per (tile, permutation chunk), operands moved by
// TMA, the permutation operand resident in shared memory.
//
// Reference work replaced: the B calls of derfinderHelper::fstats.apply inside the permutation loop of
// calculatePvalues (R/calculatePvalues.R:306-343).  Same contract as fstat_tc.cu: the tensor cores only
// SCREEN -- every candidate (base, permutation) is recomputed with the exact FP64 arithmetic of fstat.cu
// and thresholded exactly, so the output is bit-identical to the dense FP64 kernel.
//
// What changed against the first fused kernel (fstat_tc.cu), and why (profiles/r1_e_kernels.txt,
// profiles/r2_a_probe.txt):
//   * ONE fp16 product instead of three bf16 ones.  The screen only has to contain the exceedances; with
//     fp16 operands (unit round-off 2^-11) the statistic is off by <= ~1e-3 |yc| per column, which lowers
//     the candidate threshold by a fraction of a percent -- a few more candidates for the FP64 refinement,
//     a third of the tensor work and half of the operand bytes.
//   * The A operand (centred coverage as fp16, K-major rows) is written ONCE per coverage handle by
//     build_a16_kernel, together with |yc|^2 and the row mean in FP64, and loaded by TMA
//     (cp.async.bulk.tensor, 128-byte swizzle) into a 3-deep ring: no SM thread touches operand staging,
//     tile t+1 and t+2 load under the MMAs of tile t.
//   * The B operand (permuted, weighted basis rows) used to be streamed from L2 for every tile: 630 KB
//     per tile at the chr1 shape, 120 GB per launch -- more than L2 can deliver in the time the MMAs take.
//     Now a GROUP of permutation chunks stays resident in shared memory while the CTA walks all of its
//     tiles; the scan makes ceil(chunks / group) passes over the (small) fp16 A matrix instead.
//   * Three or four accumulator stages in TMEM (160 or 128 columns each) hide the MMA -> epilogue -> MMA
//     hand-off; one group of four epilogue warps (one per TMEM lane quarter) per stage scans its accumulator
//     with packed fma.rn.f32x2 on permutation pairs (the B rows are ordered so that neighbouring TMEM columns
//     belong to neighbouring permutations) and takes ONE vote per 8 permutations when nothing passes.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <math.h>
#include <stdlib.h>

#include <algorithm>

#include "data.cuh"
#include "tc_ptx.cuh"

namespace dfb {

constexpr int N2_CHUNK = 2048;  // values a warp reserves from the global cursor at a time

// ------------------------------------------------------------------------------------------
// geometry

struct Null2Geom {
    int n, npad, nk16, KB;   // samples, padded to 16, K steps of 16, 64-element K blocks per operand row
    int p, p0, j0, pe;       // model columns, first screened column, screened columns
    int pp;                  // permutations per epilogue pass (TMEM columns of a pass: [column][permutation])
    int ps, nmma;            // permutations per chunk, MMA N = ps * pe
    int nt;                  // accumulator stages
    int ew, hp;              // epilogue warps per TMEM lane quarter, permutations of a chunk per warp
    int cg;                  // chunks resident in shared memory (one group)
    int na;                  // A tile stages
    int mw;                  // MMA-issuing warps (chunks alternate between them)
    int qs, ixs;             // Qs row stride (doubles), bytes per row of the byte permutation table
    int64_t B, chunks, groups, tiles;
    uint32_t a_bytes, b_bytes;
    uint32_t off_b, off_qs, off_bar, off_warp, warp_bytes;
    size_t smem;
};

__host__ __device__ constexpr int n2_pass_perms(int pe) { return pe <= 2 ? 16 : 8; }

// ------------------------------------------------------------------------------------------
// A operand: centred log2 coverage as fp16, row-major [rows_pad][KA] (K-major rows, zero padded), plus
// the FP64 row mean and |yc|^2 of the arithmetic contract (fstat.cu).  One thread per base.

__global__ void __launch_bounds__(128) build_a16_kernel(const double* __restrict__ Y, int64_t ld, int64_t N, int n,
                                                        int KA, int center, __half* __restrict__ a16,
                                                        double* __restrict__ yy_out, double* __restrict__ m_out,
                                                        int64_t rows_pad) {
    const int64_t base = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (base >= rows_pad) return;
    uint4* dst = reinterpret_cast<uint4*>(a16 + (size_t)base * KA);
    if (base >= N) {
        for (int c = 0; c < KA / 8; ++c) dst[c] = make_uint4(0u, 0u, 0u, 0u);
        return;
    }
    double acc = 0.0;
    constexpr int LB = 8;  // independent loads in flight, then the ordered sum
    for (int k0 = 0; k0 < n; k0 += LB) {
        double v[LB];
#pragma unroll
        for (int e = 0; e < LB; ++e) v[e] = (k0 + e < n) ? Y[(size_t)(k0 + e) * ld + base] : 0.0;
#pragma unroll
        for (int e = 0; e < LB; ++e)
            if (k0 + e < n) acc = (k0 + e == 0) ? v[e] : __dadd_rn(acc, v[e]);
    }
    const double m = center ? __ddiv_rn(acc, (double)n) : 0.0;
    double yy = 0.0;
    for (int k0 = 0; k0 < KA; k0 += 8) {
        __half2 h[4];
        double raw[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) raw[e] = (k0 + e < n) ? Y[(size_t)(k0 + e) * ld + base] : 0.0;  // L2-hot second pass
#pragma unroll
        for (int e = 0; e < 8; e += 2) {
            double v[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                v[u] = 0.0;
                if (k0 + e + u < n) {
                    v[u] = __dsub_rn(raw[e + u], m);
                    yy = fma(v[u], v[u], yy);
                }
            }
            h[e >> 1] = __halves2half2(__double2half(v[0]), __double2half(v[1]));
        }
        dst[k0 >> 3] = *reinterpret_cast<const uint4*>(h);
    }
    yy_out[base] = yy;
    m_out[base] = m;
}

// B operand: per chunk a [nmma x KB*64] fp16 tile in the 128-byte swizzled UMMA layout.  Row r of a
// chunk = pass * (pp * pe) + jj * pp + qq  holds column j0 + jj of permutation  pass * pp + qq  of the
// chunk, pre-scaled by sqrt(w_j): inside an epilogue pass the TMEM columns are [column][permutation], so
// two neighbouring registers of a TMEM load are the same column of two neighbouring permutations.
__global__ void __launch_bounds__(128) build_qb16_kernel(const double* __restrict__ q /*[n][p]*/,
                                                         const int32_t* __restrict__ idx /*[B][n]*/, Null2Geom G,
                                                         double sw0, double swd, unsigned char* __restrict__ out,
                                                         unsigned char* __restrict__ idx8) {
    const int64_t chunks_per_row = (int64_t)G.KB * 8;  // 16-byte pieces of 8 consecutive K values
    const int64_t total = G.chunks * G.nmma * chunks_per_row;
    const int pass_cols = G.pp * G.pe;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = t / (G.nmma * chunks_per_row);
        const int64_t rem = t - c * G.nmma * chunks_per_row;
        const int row = (int)(rem / chunks_per_row);
        const int kc = (int)(rem - (int64_t)row * chunks_per_row);
        const int pass = row / pass_cols, within = row - pass * pass_cols;
        const int jj = within / G.pp, qq = within - jj * G.pp;
        const int j = G.j0 + jj;
        // rows beyond ps * pe pad the MMA N to a multiple of 16: zero
        const int64_t b = row < G.ps * G.pe ? c * G.ps + pass * G.pp + qq : G.B;
        __half vals[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int k = kc * 8 + e;
            double v = 0.0;
            if (b < G.B && k < G.n) v = q[(size_t)idx[b * G.n + k] * G.p + j] * (j < G.p0 ? sw0 : swd);
            vals[e] = __double2half(v);
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

// ------------------------------------------------------------------------------------------

struct Null2Params {
    const double* Y;       // log2 matrix, or nullptr when it is not materialised:
    const int32_t* xi;     //   then the integer coverage [n][ld] ...
    const double* lut;     //   ... and the log2 table (every value is inside it, checked by dfb_preprocess)
    int64_t ld, N;
    Null2Geom G;
    const unsigned char* qb;    // [chunks][KB][nmma][128 B]
    const unsigned char* idx8;  // [chunks*ps + 32][ixs]
    const double* q;            // [n][p] row-major FP64
    const double* yy;           // [N] |yc|^2 (FP64 contract)
    const double* mrow;         // [N] row means
    const float* thr;           // [N] candidate threshold of the screen statistic (null2_thr_kernel)
    double U, adjust_f, dfnum, dfden;
    uint32_t idesc;
    uint32_t* mask;  // [B][W]
    int64_t W;
    unsigned char* gflag;  // [B][Gr]: set to 1 for every 32-word group with a non-empty word
    int64_t Gr;
    int64_t* off;    // [B][W], written for non-empty words only
    double* vals;
    unsigned long long capacity;
    unsigned long long* cursor;
    unsigned long long* prof;  // [24] cycle counters (dbg & 64)
    int dbg;  // developer timing experiments (results invalid unless 64 only): 2 = one MMA per chunk, 4 = no scan,
              // 8 = no zero mask words, 64 = phase cycle counters
};

// Candidate threshold per base for one cutoff:  candidate iff  L~ >= (sqrt(T) - slack)^2  with
// T = U * (adjustF + |yc|^2 / dfden)  and  slack = kd |yc| + e0 (|m| + |yc|) + kabs (|yc| + swmax)
// (error bound of the fp16 screen, see fstat_perm_batch_null2); FP32, every rounding pushed to the safe side.
__global__ void __launch_bounds__(256) null2_thr_kernel(const double* __restrict__ yy, const double* __restrict__ mrow,
                                                        int64_t N, double U, double adjust_f, double dfden, float kd,
                                                        float e0, float kabs, float swmax, float* __restrict__ thr) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        const double y2 = yy[i];
        const float yn = __fsqrt_ru((float)y2 * 1.0000002f);
        const float Tf = (float)(U * (adjust_f + y2 / dfden));
        const float r = __fsqrt_rd(Tf) * 0.999999f - kd * yn - e0 * (fabsf((float)mrow[i]) * 1.0000002f + yn) - kabs * (yn + swmax);
        const float t2 = r > 0.f ? r * r * 0.99999f : 0.f;       // 1e-5: fp32 epilogue rounding
        thr[i] = (t2 == t2) ? t2 : __int_as_float(0x7fc00000);   // NaN data -> never a candidate
    }
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int32_t x, int32_t y, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(tm)), "r"(x), "r"(y), "r"(bar)
        : "memory");
}
__device__ __forceinline__ uint64_t pack_f32x2(uint32_t lo, uint32_t hi) {
    uint64_t v;
    asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "r"(lo), "r"(hi));
    return v;
}
__device__ __forceinline__ void unpack_f32x2(uint64_t v, float& lo, float& hi) {
    uint32_t a, b;
    asm("mov.b64 {%0, %1}, %2;" : "=r"(a), "=r"(b) : "l"(v));
    lo = __uint_as_float(a);
    hi = __uint_as_float(b);
}

// instruction descriptor, kind::f16: D = f32, A = B = f16 (format 0), both K-major, M x N
static uint32_t umma_idesc_f16(int M, int N) {
    uint32_t d = 0;
    d |= 1u << 4;
    d |= (uint32_t)(N >> 3) << 17;
    d |= (uint32_t)(M >> 4) << 24;
    return d;
}

// PT: model columns (2..8); J0: 1 when the constant first column is not screened.
// Warps: 4 * ew epilogue warps (warp w: TMEM lane quarter w & 3, share w >> 2 of a chunk's permutations),
// then the producer warp (TMA A tiles, bulk copies of the B group) and mw MMA warps (one lane each; the
// chunks alternate between them: one thread spends ~600 cycles per chunk on waits, fences, four MMA
// issues and a commit, which is more than the tensor pipe needs to execute it).
// Default chunk width: 32 permutations (a lane per permutation in the bookkeeping; at most 256 accumulator
// columns for the model sizes the kernel is compiled for).  Measured on the chr1 shape (pe = 5, c3s workload,
// gpurun_out/r2_*_sweep.txt): 32 permutations x 3 stages x 3 warp groups 2.14 ms; 24 permutations (128 columns)
// x 4 stages x 4 groups 2.35 ms; 24 x 4 stages x 2 groups with the whole accumulator row in one TMEM round trip
// 2.41 ms (two warps per scheduler cannot hide their own instruction latencies).
__host__ __device__ constexpr int n2_default_ps(int pe) {
    int ps = n2_pass_perms(pe);
    while (ps + n2_pass_perms(pe) <= 32 && ((ps + n2_pass_perms(pe)) * pe + 15) / 16 * 16 <= 256) ps += n2_pass_perms(pe);
    return ps;
}
// epilogue warp groups the kernel is compiled for = accumulator stages at the default chunk width (a group owns a
// stage): 4 for up to four screened columns, 3 for five, 2 beyond -- which leaves 104 / 136 / 184 registers per thread
__host__ __device__ constexpr int n2_max_groups(int pe) {
    const int nmma = (n2_default_ps(pe) * pe + 15) / 16 * 16;
    return 512 / nmma < 4 ? 512 / nmma : 4;
}
__host__ __device__ constexpr int n2_max_threads(int pe) { return 128 * n2_max_groups(pe) + 96; }

// registers of a TMEM load are valid after tcgen05.wait::ld; this ties their first use behind the wait for
// the compiler as well (it sees no dependence between the wait and plain arithmetic on the registers)
template <int NREG>
__device__ __forceinline__ void tmem_regs_ready(uint32_t* r) {
#pragma unroll
    for (int i = 0; i + 8 <= NREG; i += 8)
        asm volatile("" : "+r"(r[i]), "+r"(r[i + 1]), "+r"(r[i + 2]), "+r"(r[i + 3]), "+r"(r[i + 4]), "+r"(r[i + 5]),
                          "+r"(r[i + 6]), "+r"(r[i + 7]));
}
// NC consecutive columns as the fewest loads, NOT waited for
template <int NC>
__device__ __forceinline__ void tmem_ld_cols_nw(uint32_t taddr, uint32_t* r) {
    static_assert(NC % 8 == 0 && NC >= 8 && NC <= 64, "column count");
    int c = 0;
    if (NC >= 32) {
        tmem_ld32_nw(taddr, r);
        c = 32;
        if (NC == 64) {
            tmem_ld32_nw(taddr + 32u, r + 32);
            c = 64;
        }
    }
    if (NC - c >= 16) {
        tmem_ld16_nw(taddr + (uint32_t)c, r + c);
        c += 16;
    }
    if (NC - c >= 8) tmem_ld8_nw(taddr + (uint32_t)c, r + c);
}

template <int PT, int J0>
__global__ void __launch_bounds__(n2_max_threads(PT - J0), 1) fstat_null2_kernel(const Null2Params P,
                                                                        const __grid_constant__ CUtensorMap tmA) {
    constexpr int PE = PT - J0;
    constexpr int PP = n2_pass_perms(PE);
    constexpr int NC = PP * PE;  // TMEM columns per pass
    static_assert(NC % 8 == 0 && NC <= 64, "pass width");
    extern __shared__ __align__(1024) unsigned char smem_n2_raw[];
    unsigned char* sm = smem_n2_raw + ((1024u - (smem_u32(smem_n2_raw) & 1023u)) & 1023u);
    const Null2Geom& G = P.G;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n = G.n;
    const int ew = G.ew, prod_warp = 4 * G.ew, mma_warp = 4 * G.ew + 1, mw = G.mw;
    const int quarter = warp & 3, hsub = warp >> 2;  // epilogue warps only
    const int NT = G.nt, NA = G.na;
    unsigned char* As = sm;
    unsigned char* Bs = sm + G.off_b;
    double* Qs = reinterpret_cast<double*>(sm + G.off_qs);  // [n][qs]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + G.off_bar);
    // barriers: a_full[4] a_empty[4] tmem_full[4] tmem_empty[4] b_full b_empty, then the TMEM slot
    const uint32_t a_full = smem_u32(bars + 0), a_empty = smem_u32(bars + 4);
    const uint32_t tmem_full = smem_u32(bars + 8), tmem_empty = smem_u32(bars + 12);
    const uint32_t b_full = smem_u32(bars + 16), b_empty = smem_u32(bars + 17);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);
    unsigned char* wbase = sm + G.off_warp + (size_t)(warp < prod_warp ? warp : 0) * G.warp_bytes;
    uint16_t* list_w = reinterpret_cast<uint16_t*>(wbase);           // [ps * 32] (perm << 5) | base
    uint32_t* ew_w = reinterpret_cast<uint32_t*>(list_w + G.ps * 32);  // [32] exact words

    if (warp == mma_warp) {
        if (lane == 0) {
            for (int s = 0; s < 4; ++s) {
                mbar_init(a_full + 8 * s, 1);
                mbar_init(a_empty + 8 * s, (uint32_t)mw);
                mbar_init(tmem_full + 8 * s, 1);
                mbar_init(tmem_empty + 8 * s, 4u);  // the four warps of the group that scanned the stage
            }
            mbar_init(b_full, 1);
            mbar_init(b_empty, (uint32_t)mw);
            *reinterpret_cast<volatile uint32_t*>(bars + 19) = 0u;  // issue sequence number of the MMA warps
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc(smem_u32(tmem_slot), 512u);
    } else if (warp < prod_warp) {
        for (int e = tid; e < n * G.qs; e += 128 * ew) {
            const int k = e / G.qs, j = e - k * G.qs;
            Qs[e] = j < G.p ? P.q[(size_t)k * G.p + j] : 0.0;
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int64_t my_tiles = blockIdx.x < G.tiles ? (G.tiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;
    // developer instrumentation (dbg & 64): cycles per phase of one lane per role, summed over the launch
    // (counters live in shared memory: eight 64-bit registers per thread would push the epilogue into spills)
    const bool prof = (P.dbg & 64) != 0;
    unsigned long long* pc = reinterpret_cast<unsigned long long*>(bars + 20) + (warp >= mma_warp ? 8 : (warp == prod_warp ? 16 : 0));
    if (prof && tid < 24) reinterpret_cast<unsigned long long*>(bars + 20)[tid] = 0ull;
    if (prof) __syncthreads();
    long long tprev = clock64();
    auto lap = [&](int slot) {
        if (prof) {
            const long long t = clock64();
            pc[slot] += (unsigned long long)(t - tprev);
            tprev = t;
        }
    };

    if (warp == prod_warp) {
        // ===== producer: B group per pass (bulk copies), A tiles (TMA) through the ring; whole warp, one
        // elected lane issues (see the MMA warps) =====
        const bool count_me = prof && lane == 0;
        auto lap = [&](int slot) {
            if (count_me) {
                const long long t = clock64();
                pc[slot] += (unsigned long long)(t - tprev);
                tprev = t;
            }
        };
        int stage = 0;
        uint32_t aphase = 0;  // parity of the NEXT wait on a_empty[stage] is aphase ^ 1 on the first round
        for (int64_t g = 0; g < G.groups; ++g) {
            const int cgl = (int)min((int64_t)G.cg, G.chunks - g * G.cg);
            lap(0);
            if (g > 0) mbar_wait(b_empty, (uint32_t)((g - 1) & 1));  // the MMAs of the previous pass have retired
            lap(1);  // producer: wait for the previous pass to retire
            if (elect_one()) {
                mbar_expect_tx(b_full, (uint32_t)cgl * G.b_bytes);
                for (int c = 0; c < cgl; ++c)
                    bulk_g2s(smem_u32(Bs) + (uint32_t)c * G.b_bytes, P.qb + (size_t)(g * G.cg + c) * G.b_bytes, G.b_bytes,
                             b_full);
            }
            __syncwarp();
            for (int64_t i = 0; i < my_tiles; ++i) {
                const int64_t tile = blockIdx.x + i * gridDim.x;
                lap(0);
                mbar_wait(a_empty + 8 * stage, aphase ^ 1u);
                lap(2);  // producer: wait for a free A stage
                if (elect_one()) {
                    mbar_expect_tx(a_full + 8 * stage, G.a_bytes);
                    for (int kb = 0; kb < G.KB; ++kb)
                        tma_load_2d(smem_u32(As) + (uint32_t)stage * G.a_bytes + (uint32_t)kb * (TC_M * 128u), &tmA, kb * 64,
                                    (int32_t)(tile * TC_M), a_full + 8 * stage);
                }
                __syncwarp();
                if (++stage == NA) {
                    stage = 0;
                    aphase ^= 1u;
                }
            }
        }
        lap(0);
        if (count_me && P.prof)
            for (int i = 0; i < 8; ++i) atomicAdd(&P.prof[16 + i], pc[i]);
    } else if (warp >= mma_warp) {
        // ===== MMA issuers: nk16 instructions per (tile, chunk), accumulator ring of NT stages; issuer m takes
        // the chunk units with (running index) % mw == m.  The WHOLE warp runs the loop with warp-uniform
        // values and one elected lane executes the tcgen05 instructions: inside an `if (lane == 0)` region the
        // compiler cannot prove the descriptors uniform and wraps every UTCHMMA in a vote / elect / R2UR
        // broadcast loop (~130 cycles per instruction, measured) =====
        const int me = __shfl_sync(0xffffffffu, warp - mma_warp, 0);
        const bool count_me = prof && me == 0 && lane == 0;
        auto lap = [&](int slot) {
            if (count_me) {
                const long long t = clock64();
                pc[slot] += (unsigned long long)(t - tprev);
                tprev = t;
            }
        };
        int turn = 0;  // running chunk-unit index modulo mw
        // Issue order.  Accumulator stages are shared between the issuers (NT need not be a multiple of mw),
        // and an mbarrier wait only tells parities apart: an issuer running two uses of a stage ahead of the
        // other one would take the stage's INITIAL state for "released" and overwrite an accumulator that has
        // not been scanned yet.  So unit u waits for its accumulator only after unit u - 1 has got its own (a
        // sequence number in shared memory): the use before it on the same stage has then passed its wait,
        // which puts the barrier at most one phase behind.  The MMA issues of the two warps interleave.
        // (With NT a multiple of mw every stage has one issuer and none of this is needed.)
        volatile uint32_t* issue_seq = reinterpret_cast<volatile uint32_t*>(bars + 19);
        const bool use_seq = mw > 1 && (NT % mw) != 0;
        uint32_t unit = 0;
        // K step t reads columns 16 t .. 16 t + 15: 64-column block t >> 2 (one [rows x 128 B] swizzled slab
        // each), 32 bytes per step inside the block; the start-address field counts 16-byte units.
        const uint64_t da0 = umma_desc_sw128(smem_u32(As)), db0 = umma_desc_sw128(smem_u32(Bs));
        const uint64_t a_blk = (uint64_t)((TC_M * 128u) >> 4), b_blk = (uint64_t)(((uint32_t)G.nmma * 128u) >> 4);
        const int nmm = (P.dbg & 2) ? 1 : G.nk16;
        int stage = 0, s = 0;
        uint32_t aphase = 0, tphase = 0;
        for (int64_t g = 0; g < G.groups; ++g) {
            const int cgl = (int)min((int64_t)G.cg, G.chunks - g * G.cg);
            lap(0);
            mbar_wait(b_full, (uint32_t)(g & 1));
            lap(1);  // MMA warp: wait for the B group
            for (int64_t i = 0; i < my_tiles; ++i) {
                mbar_wait(a_full + 8 * stage, aphase);
                tc_fence_after();
                lap(2);  // MMA warp: wait for the A tile
                const uint64_t astage = (uint64_t)(((uint32_t)stage * G.a_bytes) >> 4);
                for (int c = 0; c < cgl; ++c) {
                    const bool mine = turn == me;
                    if (++turn == mw) turn = 0;
                    const uint32_t u = unit++;
                    if (!mine) {
                        if (++s == NT) {
                            s = 0;
                            tphase ^= 1u;
                        }
                        continue;
                    }
                    lap(0);
                    if (use_seq)
                        while (*issue_seq != u) {
                        }
                    lap(7);  // MMA warp: wait for the other issuer
                    mbar_wait_sel(tmem_empty + 8 * s, tphase ^ 1u, P.dbg & 16);
                    tc_fence_after();
                    lap(3);  // MMA warp: wait for a free accumulator
                    const uint32_t d_addr = tmem_base + (uint32_t)(s * G.nmma);
                    const uint64_t bchunk = (uint64_t)(((uint32_t)c * G.b_bytes) >> 4);
                    if (elect_one()) {
                        if (use_seq) {  // hand the sequence over before issuing
                            __threadfence_block();
                            *issue_seq = u + 1;
                        }
                        uint64_t ad = da0 + astage, bd = db0 + bchunk;
                        for (int t = 0; t < nmm; ++t) {
                            umma_f16(d_addr, ad, bd, P.idesc, t ? 1u : 0u);
                            if ((t & 3) == 3) {  // next 64-column block
                                ad += a_blk - 6;
                                bd += b_blk - 6;
                            } else {
                                ad += 2;
                                bd += 2;
                            }
                        }
                        umma_commit(tmem_full + 8 * s);  // accumulator ready for the epilogue
                        if (c + 1 == cgl) umma_commit(a_empty + 8 * stage);  // last unit of the tile in this pass
                    }
                    __syncwarp();
                    lap(4);  // MMA warp: issue + commit
                    if (++s == NT) {
                        s = 0;
                        tphase ^= 1u;
                    }
                }
                // A stage reusable once the MMAs that read it have retired: every issuer reports (count mw); the
                // one that issued the tile's last unit did so above, right behind its MMAs
                {
                    const int last_turn = (int)((unit - 1u) % (uint32_t)mw);
                    if (last_turn != me && elect_one()) umma_commit(a_empty + 8 * stage);
                    __syncwarp();
                }
                if (++stage == NA) {
                    stage = 0;
                    aphase ^= 1u;
                }
            }
            if (elect_one()) umma_commit(b_empty);  // every MMA of this pass has retired: the B group may be replaced
            __syncwarp();
        }
        lap(0);
        if (count_me && P.prof)
            for (int i = 0; i < 8; ++i) atomicAdd(&P.prof[8 + i], pc[i]);
    } else {
        // ===== epilogue + exact refinement.  Warp group h = warp >> 2 (four warps = the four TMEM lane
        // quarters) owns every chunk unit with (running index) % ew == h and scans ALL of its permutations:
        // the groups work on different accumulator stages at the same time, so a group has ew units' worth
        // of tensor time for one unit (wait, TMEM loads, scan, hand-back) instead of all sixteen warps
        // queueing on the same unit =====
        const bool count_me = prof && warp == 0 && lane == 0;
        auto lap = [&](int slot) {  // shadows the role-level helper: only one epilogue lane counts
            if (count_me) {
                const long long t = clock64();
                pc[slot] += (unsigned long long)(t - tprev);
                tprev = t;
            }
        };
        unsigned long long chunk_ptr = 0;  // warp-uniform: next free value slot of the warp's chunk
        int chunk_left = 0;
        const bool num_is_sd = P.dfnum == 1.0;
        const uint32_t lt_mask = (1u << lane) - 1u;
        const int ps = G.ps, Bi = (int)G.B;
        int s = 0, turn = 0;
        uint32_t tphase = 0;
        for (int g = 0; g < (int)G.groups; ++g) {
            const int cgl = (int)min((int64_t)G.cg, G.chunks - (int64_t)g * G.cg);
            // the candidate threshold of a row is read one tile ahead
            float thr_next = __int_as_float(0x7f800000);  // +inf: rows beyond N never pass
            {
                const int64_t b_first = (int64_t)blockIdx.x * TC_M + quarter * 32 + lane;
                if (my_tiles > 0 && b_first < P.N) thr_next = P.thr[b_first];
            }
            for (int64_t i = 0; i < my_tiles; ++i) {
                const int64_t tile = blockIdx.x + i * gridDim.x;
                const int64_t base0 = tile * TC_M;
                const int64_t gw = tile * 4 + quarter;  // global mask word of this warp's 32 bases
                const bool active = gw < P.W;
                const float thr = thr_next;
                {
                    const int64_t b_nx = base0 + quarter * 32 + lane + (int64_t)gridDim.x * TC_M;
                    thr_next = __int_as_float(0x7f800000);
                    if (i + 1 < my_tiles && b_nx < P.N) thr_next = P.thr[b_nx];
                }
                // candidates re-read their row (L2): the log2 matrix, or the integer coverage through the table
                const double* yglob = P.Y ? P.Y + base0 + quarter * 32 : nullptr;
                const int32_t* xglob = P.Y ? nullptr : P.xi + base0 + quarter * 32;
                for (int c = 0; c < cgl; ++c) {
                    const bool mine = turn == hsub;
                    if (++turn == ew) turn = 0;
                    const int s_mine = s;
                    const uint32_t ph_mine = tphase;
                    if (++s == NT) {
                        s = 0;
                        tphase ^= 1u;
                    }
                    if (!mine) continue;
                    const int b0 = (g * G.cg + c) * ps;  // first permutation of this unit
                    const int nq = min(ps, Bi - b0);
                    lap(0);
                    mbar_wait_sel(tmem_full + 8 * s_mine, ph_mine, P.dbg & 16);
                    tc_fence_after();
                    lap(2);  // epilogue: wait for the accumulator
                    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(s_mine * G.nmma);
                    int nwork = 0;
                    // one pass = PP permutations = NC accumulator columns; the load of pass k + 1 is in flight while
                    // pass k is evaluated (two register sets, selected statically)
                    auto scan_pass = [&](const uint32_t* vr, int q0) {
                        float a[PP];
#pragma unroll
                        for (int h = 0; h < PP / 2; ++h) {
                            uint64_t acc2;
                            const uint64_t v0 = pack_f32x2(vr[2 * h], vr[2 * h + 1]);
                            asm("mul.rn.f32x2 %0, %1, %1;" : "=l"(acc2) : "l"(v0));
#pragma unroll
                            for (int jj = 1; jj < PE; ++jj) {
                                const uint64_t vj = pack_f32x2(vr[jj * PP + 2 * h], vr[jj * PP + 2 * h + 1]);
                                asm("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(acc2) : "l"(vj));
                            }
                            unpack_f32x2(acc2, a[2 * h], a[2 * h + 1]);
                        }
                        float mx = a[0];
#pragma unroll
                        for (int qq = 1; qq < PP; ++qq) mx = fmaxf(mx, a[qq]);
                        if (!__any_sync(0xffffffffu, mx >= thr)) return;  // the common case for stringent cutoffs
                        const uint16_t ent0 = (uint16_t)((q0 << 5) | lane);
#pragma unroll
                        for (int qq = 0; qq < PP; ++qq) {
                            unsigned word = __ballot_sync(0xffffffffu, a[qq] >= thr);
                            if (q0 + qq >= nq) word = 0u;  // padded permutation of the last chunk
                            if (word) {
                                if ((word >> lane) & 1u) list_w[nwork + __popc(word & lt_mask)] = (uint16_t)(ent0 + (qq << 5));
                                nwork += __popc(word);
                            }
                        }
                    };
                    // Pass by pass, one register set.  Measured dead ends at the chr1 shape (c3s, gpurun_out/r2_am_sweep_c3s.txt):
                    // a second register set with the load of pass k + 1 in flight under the evaluation of pass k
                    // (2.01 -> 2.22 ms), and handing the accumulator back right after the last load instead of after
                    // the last pass (no change): the epilogue's own latency is not what bounds the stage round trip.
                    {
                        const int nscan = (P.dbg & 4) ? 0 : nq;
                        for (int q0 = 0; q0 < nscan; q0 += PP) {
                            uint32_t va[NC];
                            tmem_ld_cols_nw<NC>(taddr + (uint32_t)(q0 * PE), va);
                            tmem_ld_wait();
                            tmem_regs_ready<NC>(va);
                            scan_pass(va, q0);
                        }
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(tmem_empty + 8 * s_mine);  // accumulator free: the next MMAs overlap the refinement
                    }
                    lap(3);  // epilogue: TMEM loads + candidate scan
                    if (nwork == 0) continue;  // the mask was zeroed before the launch: nothing to write
                    ew_w[lane] = 0u;
                    if (nwork > chunk_left) {  // warp-uniform: grab a fresh chunk (the remainder is left unused)
                        unsigned long long gcur = 0;
                        if (lane == 0) gcur = atomicAdd(P.cursor, (unsigned long long)N2_CHUNK);
                        chunk_ptr = __shfl_sync(0xffffffffu, gcur, 0);
                        chunk_left = N2_CHUNK;
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
                            const int64_t cbase = base0 + quarter * 32 + bl;
                            const double yy_c = P.yy[cbase], m_c = P.mrow[cbase];  // FP64 contract values of the row
                            const uint4* ixrow = reinterpret_cast<const uint4*>(P.idx8 + (size_t)(b0 + q) * G.ixs);
                            double s0 = 0.0, sd = 0.0;
                            // S_j in k order, s0 / sd in j order: identical chains to the dense kernel
                            constexpr int QS = (PT + 1) & ~1;
                            double S[PT];
#pragma unroll
                            for (int j = 0; j < PT; ++j) S[j] = 0.0;
                            for (int k0 = 0; k0 < n; k0 += 16) {
                                const uint4 ixv = __ldg(ixrow + (k0 >> 4));  // 16 permutation indices (bytes)
                                const uint32_t ixw[4] = {ixv.x, ixv.y, ixv.z, ixv.w};
#pragma unroll
                                for (int e16 = 0; e16 < 16; ++e16) {
                                    const int k = k0 + e16;
                                    if (k < n) {
                                        const uint32_t row = (ixw[e16 >> 2] >> ((e16 & 3) * 8)) & 0xffu;
                                        const double yraw = yglob ? yglob[(size_t)k * P.ld + bl]
                                                                  : __ldg(P.lut + xglob[(size_t)k * P.ld + bl]);
                                        const double yv = __dsub_rn(yraw, m_c);
                                        const double2* qrow = reinterpret_cast<const double2*>(Qs + (size_t)row * QS);
#pragma unroll
                                        for (int j2 = 0; j2 < QS / 2; ++j2) {
                                            const double2 qq = qrow[j2];
                                            S[2 * j2] = fma(yv, qq.x, S[2 * j2]);
                                            if (2 * j2 + 1 < PT) S[2 * j2 + 1] = fma(yv, qq.y, S[2 * j2 + 1]);
                                        }
                                    }
                                }
                            }
#pragma unroll
                            for (int j = 0; j < PT; ++j) {
                                if (j < G.p0) s0 = fma(S[j], S[j], s0);
                                else sd = fma(S[j], S[j], sd);
                            }
                            double rss1 = __dsub_rn(__dsub_rn(yy_c, s0), sd);
                            if (rss1 < 0.0) rss1 = 0.0;
                            const double num = num_is_sd ? sd : __ddiv_rn(sd, P.dfnum);
                            const double den = __dadd_rn(P.adjust_f, __ddiv_rn(rss1, P.dfden));
                            f = __ddiv_rn(num, den);
                            pass = f >= P.U;
                        }
                        const unsigned bal = __ballot_sync(0xffffffffu, pass);
                        if (pass) {
                            const unsigned long long dst = run + __popc(bal & lt_mask);
                            if (dst < P.capacity) P.vals[dst] = f;
                            atomicOr(&ew_w[q], 1u << bl);
                        }
                        run += __popc(bal);
                    }
                    __syncwarp();
                    // lane = permutation again: exact word and the offset of its first value
                    const uint32_t eword = ew_w[lane];
                    const int c2 = __popc(eword);
                    int inc2 = c2;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int t = __shfl_up_sync(0xffffffffu, inc2, o);
                        if (lane >= o) inc2 += t;
                    }
                    if (lane < nq && active && eword) {
                        const int64_t b = (int64_t)b0 + lane;
                        P.mask[b * P.W + gw] = eword;
                        P.off[b * P.W + gw] = (int64_t)(chunk_ptr + (unsigned long long)(inc2 - c2));
                        P.gflag[b * P.Gr + (gw >> 5)] = 1;
                    }
                    chunk_left -= (int)(run - chunk_ptr);
                    chunk_ptr = run;
                    __syncwarp();
                    lap(5);  // epilogue: refinement + bookkeeping
                }
            }
        }
        lap(0);
        if (prof && P.prof && warp == 0 && lane == 0)
            for (int i = 0; i < 8; ++i) atomicAdd(&P.prof[i], pc[i]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == mma_warp) tmem_dealloc(tmem_base, 512u);
}

// Restore the all-zero invariant of the context's mask: zero the 32 words of every flagged group and the flag.
// A warp reads the flags of 128 groups per step.
__global__ void __launch_bounds__(256) null2_mask_cleanup_kernel(uint32_t* __restrict__ mask, unsigned char* __restrict__ gflag,
                                                                 int64_t rows, int64_t W, int64_t Gr) {
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (int64_t row = blockIdx.y; row < rows; row += gridDim.y) {
        for (int64_t k0 = ((int64_t)blockIdx.x * wpb + (threadIdx.x >> 5)) * 128; k0 < Gr; k0 += (int64_t)gridDim.x * wpb * 128) {
            unsigned f = 0u;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int64_t k = k0 + 4 * lane + j;
                if (k < Gr && gflag[row * Gr + k]) {
                    f |= 1u << j;
                    gflag[row * Gr + k] = 0;
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
                    const int64_t w = (k0 + 4 * src + j) * 32 + lane;
                    if (w < W) mask[row * W + w] = 0u;
                }
            }
        }
    }
}

int null2_release_mask(dfb_ctx* ctx, PermBatch* b) {
    if (!b || !b->mask_cached) return DFB_OK;
    b->mask_cached = false;
    if (b->mask.p != ctx->k2_mask || b->rows == 0) return DFB_OK;
    const int64_t Gr = ceil_div(b->W, 32);
    const dim3 grid((unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(Gr, 8 * 128), ceil_div((int64_t)ctx->sm_count * 8, b->rows))),
                    (unsigned)std::min<int64_t>(b->rows, 65535));
    {
        KTimer t(ctx, DFB_K_REGIONS);
        null2_mask_cleanup_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->k2_mask, ctx->k2_gflag, b->rows, b->W, Gr);
    }
    DFB_CUDA(cudaGetLastError());
    ctx->k2_clean = true;
    return DFB_OK;
}

// ------------------------------------------------------------------------------------------
// host side

static bool n2_constant_first_column(const ModelBasis& mb) {
    if (!mb.centered || mb.p0 < 1) return false;
    const double c0 = mb.q[0];
    for (int i = 1; i < mb.n; ++i)
        if (fabs(mb.q[(size_t)i * mb.p] - c0) > 1e-14) return false;
    return c0 != 0.0;
}

static bool null2_geometry(const dfb_ctx* ctx, const dfb_cov* cov, const ModelBasis& mb, int64_t B, Null2Geom* g) {
    Null2Geom G;
    memset(&G, 0, sizeof G);
    G.n = cov->n;
    if (G.n > 255 || mb.p < 2 || mb.p > 8) return false;  // permutation rows are bytes; templated column counts
    G.npad = (int)round_up(G.n, 16);
    G.nk16 = G.npad / 16;
    G.KB = (int)ceil_div(G.npad, 64);
    G.p = mb.p;
    G.p0 = mb.p0;
    G.j0 = n2_constant_first_column(mb) ? 1 : 0;
    G.pe = mb.p - G.j0;
    if (G.pe < 1) return false;
    G.pp = n2_pass_perms(G.pe);
    G.qs = (mb.p + 1) & ~1;
    G.ixs = (int)round_up(G.n, 16);
    G.B = B;
    G.tiles = ceil_div(cov->N, TC_M);
    G.a_bytes = (uint32_t)G.KB * TC_M * 128u;
    const size_t limit = std::min<size_t>(ctx->smem_optin, 227 * 1024);
    {
        const int unit = G.pp;  // a chunk is a whole number of epilogue passes
        // Chunk width (see n2_default_ps) and as many accumulator stages as the 512 TMEM columns hold: the
        // MMA -> scan -> MMA round trip of a stage (~1500 cycles) hides behind the other stages' tensor time.
        int ps = n2_default_ps(G.pe);
        if (ctx->opt_chunk_perms > 0) ps = std::max(unit, std::min(32, (ctx->opt_chunk_perms / unit) * unit));
        for (; ps >= unit; ps -= unit) {
            const int nmma = (int)round_up(ps * G.pe, 16);
            if (nmma > 256) continue;
            G.nt = std::min(4, 512 / nmma);
            if (ctx->opt_tmem_stages > 0) G.nt = std::max(1, std::min(G.nt, ctx->opt_tmem_stages));
            // epilogue warp groups: group h scans the units with index % ew == h; NT is a multiple of ew so that
            // every stage has ONE owner group (a parity wait cannot tell a stage two uses behind from a ready one)
            int ew = ctx->opt_epi_warps > 0 ? ctx->opt_epi_warps : n2_max_groups(G.pe);
            ew = std::max(1, std::min(std::min(ew, n2_max_groups(G.pe)), G.nt));
            while (G.nt % ew) --ew;
            G.ew = ew;
            G.ps = ps;
            G.hp = ps;
            G.nmma = nmma;
            G.mw = ctx->opt_mma_warps > 0 ? std::min(ctx->opt_mma_warps, 2) : 2;
            G.b_bytes = (uint32_t)G.KB * (uint32_t)nmma * 128u;
            G.chunks = ceil_div(B, ps);
            G.warp_bytes = (uint32_t)round_up(G.ps * 32 * 2 + 32 * 4, 16);
            // A ring depth: three stages unless two leave room for a larger resident group (fewer passes over
            // the fp16 coverage: at 100 samples a pass reads 256 bytes per base)
            int na_first = ctx->opt_a_stages >= 2 && ctx->opt_a_stages <= 4 ? ctx->opt_a_stages : 3;
            {
                const size_t rest0 = (size_t)round_up((int64_t)G.n * G.qs * 8, 16) + 64 * 8 + 4 * (size_t)ew * G.warp_bytes + 1024;
                auto fit = [&](int na) -> int64_t {
                    const size_t o = (size_t)na * G.a_bytes;
                    return o + rest0 + G.b_bytes > limit ? 0 : (int64_t)((limit - o - rest0) / G.b_bytes);
                };
                const int64_t c3 = std::min<int64_t>(fit(3), G.chunks), c2 = std::min<int64_t>(fit(2), G.chunks);
                if (na_first == 3 && (c3 < 1 || (c2 > c3 && ceil_div(G.chunks, c2) < ceil_div(G.chunks, std::max<int64_t>(c3, 1)))))
                    na_first = 2;
            }
            for (int na = na_first; na >= 2; --na) {
                uint32_t o = (uint32_t)na * G.a_bytes;
                const uint32_t off_b = o;
                const size_t rest = (size_t)round_up((int64_t)G.n * G.qs * 8, 16) + 64 * 8 + 4 * (size_t)ew * G.warp_bytes + 1024;
                if ((size_t)o + G.b_bytes + rest > limit) continue;
                int64_t cg = std::min<int64_t>(G.chunks, (int64_t)((limit - o - rest) / G.b_bytes));
                if (ctx->opt_b_stages > 0) cg = std::min<int64_t>(cg, ctx->opt_b_stages);
                // the bulk copies of one group complete on one mbarrier: its transaction count is 20 bits
                cg = std::min<int64_t>(cg, (int64_t)((1u << 20) - 1) / G.b_bytes);
                if (cg < 1) continue;
                G.groups = ceil_div(G.chunks, cg);
                G.cg = (int)ceil_div(G.chunks, G.groups);  // balanced passes
                G.na = na;
                G.off_b = off_b;
                o += (uint32_t)G.cg * G.b_bytes;
                G.off_qs = o;
                o += (uint32_t)round_up((int64_t)G.n * G.qs * 8, 16);
                G.off_bar = o;
                o += 64 * 8;  // 18 barriers, the TMEM slot, 24 profiling counters
                G.off_warp = o;
                o += 4u * (uint32_t)ew * G.warp_bytes;
                G.smem = (size_t)o + 1024;  // slack for the manual 1024-byte alignment
                if (G.smem > limit) continue;
                *g = G;
                return true;
            }
        }
    }
    return false;
}

bool null2_supported(const dfb_ctx* ctx, const dfb_cov* cov, const ModelBasis& mb, double lo, double hi, double adjust_f) {
    if (!ctx->opt_screen || !ctx->opt_fused || !null2_wanted(ctx, cov->n)) return false;
    if (!(hi > 0.0) || !(lo < 0.0)) return false;  // one-sided positive cutoff only
    if (cov->n - mb.p <= 0 || !(adjust_f >= 0.0)) return false;
    const double df1 = (double)(mb.p - mb.p0), df2 = (double)(cov->n - mb.p);
    if (!(sqrt(1.0 / df1 + hi / df2) < 1.0e4)) return false;  // fp16 range of the weighted basis rows
    Null2Geom g;
    return null2_geometry(ctx, cov, mb, 1, &g);
}

static PFN_cuTensorMapEncodeTiled tensor_map_encoder() {
    static PFN_cuTensorMapEncodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(p);
    }
    return fn;
}

// fp16 A operand + row constants of a coverage handle, built on first use and kept until the log2 matrix
// changes (dfb_preprocess resets it)
static int ensure_a16(dfb_ctx* ctx, const dfb_cov* cov_c, int KA, int center, int64_t rows_pad) {
    dfb_cov* cov = const_cast<dfb_cov*>(cov_c);
    if (cov->a16.p && cov->a16_center == center && cov->a16_ka == KA && cov->a16_rows == rows_pad) return DFB_OK;
    cov->a16_center = -1;
    DFB_TRY(ensure_y(ctx, cov));  // stand-alone builder (dfb_fstats was not run with this model): reads the log2 matrix
    DFB_TRY(cov->a16.alloc(ctx, (size_t)rows_pad * KA));
    DFB_TRY(cov->a16_yy.alloc(ctx, (size_t)std::max<int64_t>(cov->N, 1)));
    DFB_TRY(cov->a16_m.alloc(ctx, (size_t)std::max<int64_t>(cov->N, 1)));
    {
        KTimer t(ctx, DFB_K_FSTAT_TC, 1);
        build_a16_kernel<<<(unsigned)ceil_div(rows_pad, 128), 128, 0, ctx->stream>>>(
            cov->y.p, cov->ld, cov->N, cov->n, KA, center, reinterpret_cast<__half*>(cov->a16.p), cov->a16_yy.p,
            cov->a16_m.p, rows_pad);
    }
    DFB_CUDA(cudaGetLastError());
    cov->a16_center = center;
    cov->a16_ka = KA;
    cov->a16_rows = rows_pad;
    return DFB_OK;
}

template <int PT>
static int launch_null2(dfb_ctx* ctx, const Null2Params& F, const CUtensorMap& tm, int grid) {
    auto go = [&](auto kern) -> int {
        DFB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)F.G.smem));
        KTimer t(ctx, DFB_K_FSTAT_TC, 1);
        kern<<<grid, 128 * F.G.ew + 32 + 32 * F.G.mw, F.G.smem, ctx->stream>>>(F, tm);
        return DFB_OK;
    };
    if (F.G.j0) return go(fstat_null2_kernel<PT, 1>);
    return go(fstat_null2_kernel<PT, 0>);
}

// Screen + refine for one batch of permutations; fills the same PermBatch as the other paths.
int fstat_perm_batch_null2(dfb_ctx* ctx, const dfb_cov* cov, const ModelBasis& mb, const double* q_dev /*[n][p]*/,
                           double adjust_f, const int32_t* idx_dev, int64_t B, double hi, size_t capacity_hint,
                           PermBatch* out, bool defer_check) {
    Null2Geom G;
    if (!null2_geometry(ctx, cov, mb, B, &G)) {
        set_error("null-scan geometry does not fit (internal error)");
        return DFB_EINVAL;
    }
    PFN_cuTensorMapEncodeTiled encode = tensor_map_encoder();
    DFB_REQUIRE(encode != nullptr, "cuTensorMapEncodeTiled is not available from the CUDA driver");
    const int KA = G.KB * 64;
    const int64_t rows_pad = G.tiles * TC_M;
    DFB_TRY(ensure_a16(ctx, cov, KA, mb.centered ? 1 : 0, rows_pad));
    CUtensorMap tm;
    {
        const cuuint64_t gdim[2] = {(cuuint64_t)KA, (cuuint64_t)rows_pad};
        const cuuint64_t gstride[1] = {(cuuint64_t)KA * 2};
        const cuuint32_t box[2] = {64, (cuuint32_t)TC_M};
        const cuuint32_t estr[2] = {1, 1};
        const CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)cov->a16.p, gdim, gstride, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        DFB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    }
    DevBuf<unsigned char> qb, idx8;
    DFB_TRY(qb.alloc(ctx, (size_t)G.chunks * G.b_bytes));
    DFB_TRY(idx8.alloc(ctx, (size_t)(G.chunks * G.ps + 32) * G.ixs));
    const double df1 = (double)(mb.p - mb.p0), df2 = (double)(cov->n - mb.p);
    const double w_a = 1.0 / df1 + hi / df2, w_c = hi / df2;  // weights of the S_d / S_0 columns
    {
        KTimer t(ctx, DFB_K_FSTAT_TC, 1);
        const int64_t work = G.chunks * G.nmma * (G.KB * 8);
        build_qb16_kernel<<<grid_for(work, 128, ctx->sm_count, 16), 128, 0, ctx->stream>>>(q_dev, idx_dev, G, sqrt(w_c),
                                                                                         sqrt(w_a), qb.p, idx8.p);
    }
    Null2Params F;
    memset(&F, 0, sizeof F);
    F.Y = cov->y.p;
    if (!F.Y) {
        DFB_REQUIRE(cov->y_lazy, "coverageProcessed is missing: run dfb_preprocess first");
        DFB_TRY(ensure_lut(ctx, cov->scalefac));
        F.xi = cov->xi.p;
        F.lut = ctx->lut_dev;
    }
    F.ld = cov->ld;
    F.N = cov->N;
    F.G = G;
    F.qb = qb.p;
    F.idx8 = idx8.p;
    F.q = q_dev;
    F.yy = cov->a16_yy.p;
    F.mrow = cov->a16_m.p;
    F.U = hi;
    F.adjust_f = adjust_f;
    F.dfnum = df1;
    F.dfden = df2;
    // Error bound of the screen.  Operands are rounded to fp16 (unit round-off 2^-11 in the normal range,
    // absolute 2^-25 below 2^-14): |a~ b~ - a b| summed over k is at most (2^-10 + 2^-22) |yc| sqrt(w_j)
    // (Cauchy-Schwarz, |q_j| = 1) plus sqrt(n) 2^-25 (|yc| + sqrt(w_j)) from the subnormal range; products of
    // two fp16 values are exact in FP32 and the accumulation of n of them adds at most n 2^-22 |yc| sqrt(w_j)
    // (truncating accumulate).  Over the screened columns, with kappa^2 = sum w_j, the computed vector is
    // within  kappa * delta_rel * |yc| + kabs * (|yc| + swmax)  of the exact one.  2x safety on both terms.
    const double kappa = sqrt(w_a * (double)(mb.p - mb.p0) + w_c * (double)mb.p0);
    const double delta_rel = 2.0 * (9.765625e-4 /*2^-10*/ + 2.3842e-7 /*2^-22*/ * (1.0 + (double)cov->n));
    const float kd = (float)(kappa * delta_rel * 1.000001);
    const float kabs = (float)(2.0 * sqrt((double)G.pe * (double)cov->n) * 2.98023223876953125e-8 /*2^-25*/ * 1.000001);
    const float swmax = (float)(sqrt(std::max(w_a, w_c)) * 1.000001);
    // dropped constant column: |S_const| = |sum_k yc[k]| / sqrt(n) <= n^1.5 * 2^-52 * max|y| (rounding of the
    // mean and of the n subtractions); 2x safety.  Its weight in the statistic is c.
    const float e0 = G.j0 ? (float)(sqrt(w_c) * 2.0 * pow((double)cov->n, 1.5) * 2.220446049250313e-16 * 1.000001) : 0.f;
    DevBuf<float> thr;
    DFB_TRY(thr.alloc(ctx, (size_t)std::max<int64_t>(cov->N, 1)));
    {
        KTimer t(ctx, DFB_K_FSTAT_TC, 1);
        null2_thr_kernel<<<grid_for(cov->N, 256, ctx->sm_count, 8), 256, 0, ctx->stream>>>(
            cov->a16_yy.p, cov->a16_m.p, cov->N, hi, adjust_f, df2, kd, e0, kabs, swmax, thr.p);
    }
    F.thr = thr.p;
    F.idesc = umma_idesc_f16(TC_M, G.nmma);
    F.dbg = ctx->opt_dbg;
    DevBuf<unsigned long long> profbuf;
    if (F.dbg & 64) {
        DFB_TRY(profbuf.alloc(ctx, 24));
        DFB_TRY(profbuf.zero());
        F.prof = profbuf.p;
    }

    out->rows = B;
    out->W = ceil_div(cov->N, 32);
    out->tile_bases = 32;
    out->tiles = out->W;
    F.Gr = ceil_div(out->W, 32);
    const size_t need_w = (size_t)out->rows * out->W, need_g = (size_t)out->rows * (size_t)F.Gr;
    const bool cached = ctx->opt_mask_cache != 0;
    if (cached) {
        // the context's kept-zero buffers (grown when a larger batch comes along)
        if (ctx->k2_mask_cap < need_w || ctx->k2_gflag_cap < need_g) {
            DFB_CUDA(cudaStreamSynchronize(ctx->stream));
            if (ctx->k2_mask) cudaFree(ctx->k2_mask);
            if (ctx->k2_gflag) cudaFree(ctx->k2_gflag);
            ctx->k2_mask = nullptr;
            ctx->k2_gflag = nullptr;
            ctx->k2_mask_cap = ctx->k2_gflag_cap = 0;
            ctx->k2_clean = false;
            if (cudaMalloc((void**)&ctx->k2_mask, need_w * sizeof(uint32_t)) != cudaSuccess ||
                cudaMalloc((void**)&ctx->k2_gflag, need_g) != cudaSuccess) {
                cudaGetLastError();
                if (ctx->k2_mask) cudaFree(ctx->k2_mask);
                ctx->k2_mask = nullptr;
                ctx->k2_gflag = nullptr;
                set_error("device allocation of the null-scan mask (%zu bytes) failed", need_w * sizeof(uint32_t));
                return DFB_ENOMEM;
            }
            ctx->k2_mask_cap = need_w;
            ctx->k2_gflag_cap = need_g;
        }
        out->mask.borrow(ctx->stream, ctx->k2_mask, need_w);
        out->gflag.borrow(ctx->stream, ctx->k2_gflag, need_g);
        out->mask_cached = true;
    } else {
        DFB_TRY(out->mask.alloc(ctx, need_w));
        DFB_TRY(out->gflag.alloc(ctx, need_g));
    }
    DFB_TRY(out->off.alloc(ctx, (size_t)out->rows * out->tiles));
    F.gflag = out->gflag.p;
    DevBuf<unsigned long long> cursor;
    DFB_TRY(cursor.alloc(ctx, 1));
    // One CTA per SM (the accumulator ring takes all 512 TMEM columns), `k2_waves` CTAs per SM slot over the
    // launch so that retiring CTAs free slots for the observed-region chain on the auxiliary stream.
    const int waves = std::max(1, std::min(ctx->opt_k2_waves, 4));
    const int grid = (int)std::min<int64_t>(G.tiles, (int64_t)ctx->sm_count * waves);
    if (getenv("DFB_GAPS"))
        fprintf(stderr,
                "[dfb] null2 geometry: n=%d KB=%d pe=%d pp=%d ps=%d nmma=%d nt=%d ew=%d mw=%d na=%d cg=%d groups=%lld chunks=%lld smem=%zu grid=%d\n",
                G.n, G.KB, G.pe, G.pp, G.ps, G.nmma, G.nt, G.ew, G.mw, G.na, G.cg, (long long)G.groups, (long long)G.chunks, G.smem, grid);
    const size_t slack = (size_t)grid * 4 * G.ew * N2_CHUNK;  // every warp may leave one partly used chunk behind
    const size_t dense = (size_t)B * (size_t)cov->N;
    size_t cap = std::min(dense, std::max<size_t>(capacity_hint, 1024)) + slack;
    F.mask = out->mask.p;
    F.W = out->W;
    F.off = out->off.p;
    F.cursor = cursor.p;
    for (int attempt = 0; attempt < 2; ++attempt) {
        DFB_TRY(out->vals.alloc(ctx, cap));
        DFB_TRY(cursor.zero());
        // the kernel writes the non-empty mask words only (at stringent cutoffs nearly all are empty, and 16
        // warps storing 4-byte zeros to 32 different rows per chunk cost more than the scan itself)
        if (!cached) {
            DFB_TRY(out->mask.zero());
            DFB_TRY(out->gflag.zero());
        } else {
            if (!ctx->k2_clean) {  // first use, or a previous user did not hand the buffers back clean
                DFB_CUDA(cudaMemsetAsync(ctx->k2_mask, 0, ctx->k2_mask_cap * sizeof(uint32_t), ctx->stream));
                DFB_CUDA(cudaMemsetAsync(ctx->k2_gflag, 0, ctx->k2_gflag_cap, ctx->stream));
            }
            ctx->k2_clean = false;  // about to be written; null2_release_mask() restores the invariant
        }
        F.vals = out->vals.p;
        F.capacity = cap;
        switch (mb.p) {
            case 2: DFB_TRY(launch_null2<2>(ctx, F, tm, grid)); break;
            case 3: DFB_TRY(launch_null2<3>(ctx, F, tm, grid)); break;
            case 4: DFB_TRY(launch_null2<4>(ctx, F, tm, grid)); break;
            case 5: DFB_TRY(launch_null2<5>(ctx, F, tm, grid)); break;
            case 6: DFB_TRY(launch_null2<6>(ctx, F, tm, grid)); break;
            case 7: DFB_TRY(launch_null2<7>(ctx, F, tm, grid)); break;
            case 8: DFB_TRY(launch_null2<8>(ctx, F, tm, grid)); break;
            default: set_error("null scan: unsupported column count %d", mb.p); return DFB_EINVAL;
        }
        DFB_CUDA(cudaGetLastError());
        if ((F.dbg & 64) && F.prof) {
            unsigned long long h[24];
            DFB_TRY(d2h(ctx, h, F.prof, sizeof h));
            const double g = (double)grid;
            fprintf(stderr, "[dfb] null2 phases, mean cycles per CTA -- epilogue warp 0: thresholds %.0f  wait-acc %.0f  scan %.0f  refine %.0f  other %.0f\n",
                    h[1] / g, h[2] / g, h[3] / g, h[5] / g, h[0] / g);
            fprintf(stderr, "[dfb] null2 phases, mean cycles per CTA -- MMA lane: wait-B %.0f  wait-A %.0f  wait-turn %.0f  wait-acc %.0f  issue %.0f  hand-over %.0f  commit %.0f  other %.0f;  producer: wait-pass %.0f  wait-A-stage %.0f  other %.0f\n",
                    h[8 + 1] / g, h[8 + 2] / g, h[8 + 7] / g, h[8 + 3] / g, h[8 + 4] / g, h[8 + 5] / g, h[8 + 6] / g, h[8 + 0] / g, h[16 + 1] / g, h[16 + 2] / g, h[16 + 0] / g);
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
