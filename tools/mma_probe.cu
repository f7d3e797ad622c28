// Developer probe (not part of the library): tcgen05.mma issue rate by N, and TMEM read rate by the
// number of epilogue warps, on one CTA per SM.  Build: nvcc -gencode arch=compute_100a,code=sm_100a
// -O3 -o tools/mma_probe tools/mma_probe.cu ; run on a B200.
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(bar),
        "r"(parity), "r"(0x989680u)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d),
                 "l"(a), "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
static uint32_t idesc_f16(int M, int N) {
    uint32_t d = 0;
    d |= 1u << 4;  // D f32, A = B = f16 (format 0)
    d |= (uint32_t)(N >> 3) << 17;
    d |= (uint32_t)(M >> 4) << 24;
    return d;
}
__device__ __forceinline__ void tmem_ld32_nw(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}

// mode 0: MMA rate. warp 0 lane 0 issues `reps` chunks of `kmma` MMAs (N columns) alternating two
// accumulator stages; it waits for the commit of chunk i-2 before issuing chunk i.
// mode 1: TMEM read rate: `nwarps` warps (quarter = warp & 3) each read `reps` x 32-column loads, summing
// squares (the screen's epilogue arithmetic) so the loads are not dead.
struct Args {
    int mode, N, kmma, reps, nwarps, ffma2;
    uint32_t idesc;
    unsigned long long* cycles;  // [grid]
    float* sink;
};

__global__ void __launch_bounds__(544) probe_kernel(Args a) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bars[4];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // A: 128 x 64 fp16 (16 KB), B: 256 x 64 fp16 (32 KB); small finite values
    for (int i = tid; i < (16 + 32) * 1024 / 2; i += blockDim.x) reinterpret_cast<__half*>(sm)[i] = __float2half(0.001f * (i & 63));
    if (warp == 0) {
        if (lane == 0) {
            mbar_init(smem_u32(&bars[0]), 1);
            mbar_init(smem_u32(&bars[1]), 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    long long t0 = clock64(), t1 = t0;
    if (a.mode == 0) {
        if (warp == 0 && lane == 0) {
            const uint32_t a_addr = smem_u32(sm), b_addr = smem_u32(sm + 16 * 1024);
            uint32_t par[2] = {0, 0};
            t0 = clock64();
            for (int i = 0; i < a.reps; ++i) {
                const int s = i & 1;
                if (i >= 2) {
                    mbar_wait(smem_u32(&bars[s]), par[s]);
                    par[s] ^= 1;
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                for (int k = 0; k < a.kmma; ++k)
                    umma(tmem + (uint32_t)(s * 256), desc_sw128(a_addr + (k & 3) * 32u), desc_sw128(b_addr + (k & 3) * 32u), a.idesc,
                         k ? 1u : 0u);
                umma_commit(smem_u32(&bars[s]));
            }
            for (int s = 0; s < 2; ++s) {
                const int last = ((a.reps - 1 - s) & 1);
                mbar_wait(smem_u32(&bars[last]), par[last]);
                par[last] ^= 1;
            }
            t1 = clock64();
            a.cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
        }
    } else {
        if (warp < a.nwarps) {
            const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
            float acc = 0.f;
            const float thr = 1e30f;
            unsigned any = 0;
            t0 = clock64();
            for (int i = 0; i < a.reps; ++i) {
                uint32_t r[32];
                tmem_ld32_nw(taddr + (uint32_t)((i * 32) & 255), r);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (a.ffma2) {
                    // 5 columns x 6 permutations laid out [j][q]: pairs of permutations per packed FMA
                    unsigned long long s2[3];
#pragma unroll
                    for (int q2 = 0; q2 < 3; ++q2) {
                        unsigned long long v = ((unsigned long long)r[2 * q2 + 1] << 32) | r[2 * q2];
                        asm("mul.f32x2 %0, %1, %1;" : "=l"(s2[q2]) : "l"(v));
#pragma unroll
                        for (int j = 1; j < 5; ++j) {
                            unsigned long long w = ((unsigned long long)r[j * 6 + 2 * q2 + 1] << 32) | r[j * 6 + 2 * q2];
                            asm("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(s2[q2]) : "l"(w));
                        }
                    }
                    float m = 0.f;
#pragma unroll
                    for (int q2 = 0; q2 < 3; ++q2) {
                        m = fmaxf(m, __uint_as_float((uint32_t)s2[q2]));
                        m = fmaxf(m, __uint_as_float((uint32_t)(s2[q2] >> 32)));
                    }
                    any |= __ballot_sync(0xffffffffu, m >= thr);
                    acc += m;
                } else {
#pragma unroll
                    for (int q = 0; q < 6; ++q) {
                        float s = __uint_as_float(r[q * 5]) * __uint_as_float(r[q * 5]);
#pragma unroll
                        for (int j = 1; j < 5; ++j) s = fmaf(__uint_as_float(r[q * 5 + j]), __uint_as_float(r[q * 5 + j]), s);
                        any |= __ballot_sync(0xffffffffu, s >= thr);
                        acc += s;
                    }
                }
            }
            t1 = clock64();
            if (lane == 0 && warp == 0) a.cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
            if (acc == 123.f || any) a.sink[0] = acc;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
    int dev = 0, sms = 0;
    cudaSetDevice(dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    unsigned long long* cyc;
    float* sink;
    cudaMalloc(&cyc, sizeof(unsigned long long) * sms);
    cudaMalloc(&sink, 4);
    const size_t smem = 49 * 1024 + 1024;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    unsigned long long* h = (unsigned long long*)malloc(sizeof(unsigned long long) * sms);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    auto run = [&](Args a, int threads, const char* label) {
        a.cycles = cyc;
        a.sink = sink;
        cudaMemset(cyc, 0, sizeof(unsigned long long) * sms);
        probe_kernel<<<sms, threads, smem>>>(a);  // warm-up
        cudaEventRecord(e0);
        probe_kernel<<<sms, threads, smem>>>(a);
        cudaEventRecord(e1);
        cudaError_t err = cudaDeviceSynchronize();
        if (err != cudaSuccess) {
            printf("%s: CUDA error %s\n", label, cudaGetErrorString(err));
            exit(1);
        }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        cudaMemcpy(h, cyc, sizeof(unsigned long long) * sms, cudaMemcpyDeviceToHost);
        double mean = 0;
        for (int i = 0; i < sms; ++i) mean += (double)h[i];
        mean /= sms;
        printf("%s: %.1f cycles/unit (kernel %.3f ms, %.0f MHz implied)\n", label, mean / ((double)a.reps * (a.mode == 0 ? a.kmma : 1)), ms,
               mean / (ms * 1e3));
    };
    const int Ns[] = {64, 80, 96, 128, 160, 200, 240, 256};
    for (int N : Ns)
        for (int kmma : {4, 8, 12}) {
            Args a = {};
            a.mode = 0;
            a.N = N;
            a.kmma = kmma;
            a.reps = 2000;
            a.idesc = idesc_f16(128, N);
            char label[128];
            snprintf(label, sizeof label, "mma M=128 N=%d K=16 x%d per chunk", N, kmma);
            run(a, 32, label);
        }
    for (int ffma2 : {0, 1})
        for (int nw : {4, 8, 12, 16}) {
            Args a = {};
            a.mode = 1;
            a.reps = 4000;
            a.nwarps = nw;
            a.ffma2 = ffma2;
            char label[128];
            snprintf(label, sizeof label, "tmem 32-col load + 6-perm x 5-col stat, %d warps, ffma2=%d (per load per warp)", nw, ffma2);
            run(a, nw * 32, label);
        }
    return 0;
}
