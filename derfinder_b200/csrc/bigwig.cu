// BigWig data-section decode on the device (R/railMatrix.R:271-289).
//
// railMatrix() imports, per sample, the ranges of the expressed regions from the sample's BigWig
// (`import(sampleFile, selection = regs, as = "RleList")`), divides by mappedPerXM and takes
// sum(Views(cov, regs)) / L.  File access, the R-tree look-up and zlib inflation stay on the host (rtracklayer
// in R, derfinder_b200/bigwig.py in the Python mirror); the inflated data sections -- the bulk of the bytes --
// are shipped as they are and decoded here: every section is a 24-byte header + bedGraph / variableStep /
// fixedStep items (UCSC bigWig format, little endian).  The decode writes value / mappedPerXM of every base
// that falls inside a region into a dense [n][sum(width(regions))] double matrix (zero where the file has no
// data, as rtracklayer's RleList has), which dfb_view_sums then reduces in IRanges' order.
#include "data.cuh"

namespace dfb {

struct BwSectionHeader {
    uint32_t chrom_id, chrom_start, chrom_end, item_step, item_span;
    uint8_t type, reserved;
    uint16_t item_count;
};
static_assert(sizeof(BwSectionHeader) == 24, "bigWig section header");

// first region whose end is >= pos (1-based), R when none
__device__ __forceinline__ int64_t bw_first_region(const int32_t* __restrict__ rend, int64_t R, int64_t pos) {
    int64_t lo = 0, hi = R;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if ((int64_t)rend[mid] < pos) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// one block per section; a thread per item, which writes its overlap with the (sorted, disjoint) regions
__global__ void __launch_bounds__(128) bigwig_decode_kernel(const unsigned char* __restrict__ blob,
                                                            const int64_t* __restrict__ sec_off, int64_t nsec,
                                                            uint32_t chrom_id, const int32_t* __restrict__ rstart,
                                                            const int32_t* __restrict__ rend,
                                                            const int64_t* __restrict__ roff, int64_t R, double divisor,
                                                            double* __restrict__ out, int* __restrict__ bad) {
    for (int64_t sec = blockIdx.x; sec < nsec; sec += gridDim.x) {
        const unsigned char* p = blob + sec_off[sec];
        const int64_t bytes = sec_off[sec + 1] - sec_off[sec];
        if (bytes < 24) {
            if (threadIdx.x == 0) atomicExch(bad, 1);
            continue;
        }
        BwSectionHeader h;
        memcpy(&h, p, sizeof h);  // sections are not aligned inside the blob
        const int item_bytes = h.type == 1 ? 12 : (h.type == 2 ? 8 : 4);
        if (h.type < 1 || h.type > 3 || 24 + (int64_t)h.item_count * item_bytes > bytes) {
            if (threadIdx.x == 0) atomicExch(bad, 2);
            continue;
        }
        if (h.chrom_id != chrom_id) continue;  // a leaf block may span chromosomes
        for (int it = threadIdx.x; it < (int)h.item_count; it += blockDim.x) {
            const unsigned char* q = p + 24 + (size_t)it * item_bytes;
            uint32_t s0, e0;
            float val;
            if (h.type == 1) {  // bedGraph: start, end (0-based half open), value
                memcpy(&s0, q, 4);
                memcpy(&e0, q + 4, 4);
                memcpy(&val, q + 8, 4);
            } else if (h.type == 2) {  // variableStep: start, value; span from the header
                memcpy(&s0, q, 4);
                memcpy(&val, q + 4, 4);
                e0 = s0 + h.item_span;
            } else {  // fixedStep: value; start from the header
                memcpy(&val, q, 4);
                s0 = h.chrom_start + (uint32_t)it * h.item_step;
                e0 = s0 + h.item_span;
            }
            const int64_t a = (int64_t)s0 + 1, b = (int64_t)e0;  // 1-based inclusive
            if (b < a) continue;
            const double v = __ddiv_rn((double)val, divisor);
            for (int64_t k = bw_first_region(rend, R, a); k < R && (int64_t)rstart[k] <= b; ++k) {
                const int64_t lo = max(a, (int64_t)rstart[k]), hi = min(b, (int64_t)rend[k]);
                double* dst = out + roff[k] + (lo - rstart[k]);
                for (int64_t i = 0; i <= hi - lo; ++i) dst[i] = v;
            }
        }
    }
}

}  // namespace dfb

using namespace dfb;

extern "C" int dfb_cov_from_bigwig(dfb_ctx* ctx, int n, const void* const* sections, const int64_t* const* sec_off,
                                   const int64_t* nsec, const uint32_t* chrom_id, const int32_t* rstart,
                                   const int32_t* rend, int64_t R, const double* mapped_per_xm, dfb_cov** out) {
    DFB_REQUIRE(ctx && out && n > 0 && sections && sec_off && nsec && chrom_id, "dfb_cov_from_bigwig: NULL argument");
    DFB_REQUIRE(R > 0 && rstart && rend, "dfb_cov_from_bigwig: no regions");
    *out = nullptr;
    std::vector<int64_t> roff((size_t)R + 1, 0);
    std::vector<int32_t> pos_len;
    int pos_first = 0;
    for (int64_t k = 0; k < R; ++k) {
        DFB_REQUIRE(rstart[k] >= 1 && rend[k] >= rstart[k], "region %lld: bad range [%d, %d]", (long long)k + 1, rstart[k], rend[k]);
        DFB_REQUIRE(k == 0 || rstart[k] > rend[k - 1], "regions must be sorted and disjoint (region %lld)", (long long)k + 1);
        roff[(size_t)k + 1] = roff[(size_t)k] + ((int64_t)rend[k] - rstart[k] + 1);
        const int32_t gap = k == 0 ? rstart[0] - 1 : rstart[k] - rend[k - 1] - 1;
        if (k == 0 && gap == 0) pos_first = 1;
        if (gap > 0) pos_len.push_back(gap);
        if (k > 0 && gap == 0) pos_len.back() += rend[k] - rstart[k] + 1;  // abutting regions: one TRUE run
        else pos_len.push_back(rend[k] - rstart[k] + 1);
    }
    const int64_t N = roff[(size_t)R];
    dfb_cov* c = new dfb_cov();
    c->ctx = ctx;
    c->n = n;
    c->N = N;
    c->dtype = DFB_F64;
    c->ld = round_up(std::max<int64_t>(N, 1), 64);
    struct Guard {
        dfb_cov* p;
        ~Guard() {
            if (p) {
                cudaStreamSynchronize(p->ctx->stream);
                delete p;
            }
        }
    } guard{c};
    DFB_TRY(c->xd.alloc(ctx, (size_t)n * c->ld));
    DFB_TRY(c->xd.zero());  // bases without data are 0 (rtracklayer's RleList)
    DFB_TRY(c->pos.build(ctx, pos_len.data(), (int64_t)pos_len.size(), pos_first));
    DevBuf<int32_t> rs, re;
    DevBuf<int64_t> ro;
    DevBuf<int> bad;
    DFB_TRY(rs.alloc(ctx, (size_t)R));
    DFB_TRY(re.alloc(ctx, (size_t)R));
    DFB_TRY(ro.alloc(ctx, (size_t)R + 1));
    DFB_TRY(bad.alloc(ctx, 1));
    DFB_TRY(bad.zero());
    DFB_TRY(h2d(ctx, rs.p, rstart, (size_t)R * 4));
    DFB_TRY(h2d(ctx, re.p, rend, (size_t)R * 4));
    DFB_TRY(h2d(ctx, ro.p, roff.data(), roff.size() * 8));
    for (int s = 0; s < n; ++s) {
        const int64_t ns = nsec[s];
        if (ns == 0) continue;
        DFB_REQUIRE(sections[s] && sec_off[s], "dfb_cov_from_bigwig: sample %d has sections but no data", s + 1);
        const int64_t bytes = sec_off[s][ns];
        DevBuf<unsigned char> blob;
        DevBuf<int64_t> off;
        DFB_TRY(blob.alloc(ctx, (size_t)std::max<int64_t>(bytes, 1)));
        DFB_TRY(off.alloc(ctx, (size_t)ns + 1));
        DFB_TRY(h2d_big(ctx, blob.p, sections[s], (size_t)bytes));
        DFB_TRY(h2d(ctx, off.p, sec_off[s], (size_t)(ns + 1) * 8));
        const double div = mapped_per_xm ? mapped_per_xm[s] : 1.0;
        {
            KTimer t(ctx, DFB_K_REGMAT);
            bigwig_decode_kernel<<<(unsigned)std::min<int64_t>(ns, (int64_t)ctx->sm_count * 16), 128, 0, ctx->stream>>>(
                blob.p, off.p, ns, chrom_id[s], rs.p, re.p, ro.p, R, div, c->xd.p + (size_t)s * c->ld, bad.p);
        }
        DFB_CUDA(cudaGetLastError());
        DFB_CUDA(cudaStreamSynchronize(ctx->stream));  // the staging buffers of this sample die here
    }
    int hbad = 0;
    DFB_TRY(d2h(ctx, &hbad, bad.p, sizeof hbad));
    DFB_REQUIRE(hbad == 0, "dfb_cov_from_bigwig: malformed data section (%s)", hbad == 1 ? "shorter than its header" : "bad type or item count");
    guard.p = nullptr;
    *out = c;
    return DFB_OK;
}
