// K5: posterior accumulators straight from device-resident sample records -- the per-bin marginal rates, shift-time
// histograms and number-of-rates counts that plotRJforward.v3.py derives from the text logs one row at a time
// (get_marginal_rates :92-139, get_r_plot :166-176, get_K_values :292-305).  SURVEY 8 f-1: for ensembles of thousands of
// chains the records never have to leave the GPU (nor be formatted as text) to be summarised.
//
//   marginal rate of bin j in one sample = rates[ number of shift times falling in bins 0..j ]      (np.histogram semantics:
//   half-open unit bins from `first_edge`, the last one closed, times outside ignored)
//
// One warp per record (grid-stride, fixed grid => fixed summation order => deterministic): lane k holds rate k and shift k,
// lanes own bins j = lane, lane+32, ...; partial sums stay in registers across records and go to [warp][...] partials that a
// second kernel adds in order.  HBM-bound: 1152 B per record.
#include "lr_common.cuh"

namespace {

constexpr int K5_BINS_PER_LANE = 8;                 // 256 bins per pass
constexpr int K5_WARPS_PER_CTA = 8;
// record layout (include/literate_b200.h)
constexpr int REC_KL = 5, REC_KM = 6, REC_L = 16, REC_TL = 48, REC_M = 80, REC_TM = 112, REC_W = LR_REC_DOUBLES;

struct K5Params {
    const double* rec;
    long long n_rec;            // records to use (already past the burn-in)
    double first_edge;
    int nb;
    int bin0;                   // first bin of this pass
    double* part_rate;          // [n_warps][2][256]
    long long* part_cnt;        // [n_warps][2][256 + 32]   shift counts, then K counts
};

__device__ __forceinline__ void k5_side(const double* r, int off_k, int off_r, int off_t, double e0, int nb, int bin0, int lane,
                                        double (&acc)[K5_BINS_PER_LANE], int (&cnt)[K5_BINS_PER_LANE], int& kcnt) {
    const int K = (int)r[off_k];
    const double rate = lane < K ? r[off_r + lane] : 0.0;
    // slot k >= 1 holds shift k-1; its histogram bin (np.histogram: half-open unit bins, the last one closed, outside ignored)
    int sb = 0x7fffffff;
    if (lane >= 1 && lane < K) {
        const double s = r[off_t + lane];
        if (s >= e0 && s <= e0 + (double)nb) { sb = __double2int_rd(s - e0); if (sb > nb - 1) sb = nb - 1; }
    }
    if (bin0 == 0 && lane == K - 1) kcnt++;
    int idx[K5_BINS_PER_LANE], here[K5_BINS_PER_LANE];
#pragma unroll
    for (int q = 0; q < K5_BINS_PER_LANE; ++q) { idx[q] = 0; here[q] = 0; }
    for (int k = 1; k < K; ++k) {
        const int b = __shfl_sync(0xffffffffu, sb, k);
#pragma unroll
        for (int q = 0; q < K5_BINS_PER_LANE; ++q) {
            const int j = bin0 + lane + 32 * q;
            idx[q] += b <= j ? 1 : 0;
            here[q] += b == j ? 1 : 0;
        }
    }
#pragma unroll
    for (int q = 0; q < K5_BINS_PER_LANE; ++q) {
        const int j = bin0 + lane + 32 * q;
        const double v = __shfl_sync(0xffffffffu, rate, idx[q] & 31);
        if (j < nb) { acc[q] += v; cnt[q] += here[q]; }
    }
}

__global__ void __launch_bounds__(K5_WARPS_PER_CTA * 32) k5_accumulate_kernel(const K5Params p) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    double accL[K5_BINS_PER_LANE], accM[K5_BINS_PER_LANE];
    int cntL[K5_BINS_PER_LANE], cntM[K5_BINS_PER_LANE];
    int kL = 0, kM = 0;
#pragma unroll
    for (int q = 0; q < K5_BINS_PER_LANE; ++q) { accL[q] = 0.0; accM[q] = 0.0; cntL[q] = 0; cntM[q] = 0; }
    for (long long i = warp; i < p.n_rec; i += n_warps) {
        const double* r = p.rec + (size_t)i * REC_W;
        k5_side(r, REC_KL, REC_L, REC_TL, p.first_edge, p.nb, p.bin0, lane, accL, cntL, kL);
        k5_side(r, REC_KM, REC_M, REC_TM, p.first_edge, p.nb, p.bin0, lane, accM, cntM, kM);
    }
    double* pr = p.part_rate + (size_t)warp * 2 * 256;
    long long* pc = p.part_cnt + (size_t)warp * 2 * (256 + 32);
#pragma unroll
    for (int q = 0; q < K5_BINS_PER_LANE; ++q) {
        pr[lane + 32 * q] = accL[q]; pr[256 + lane + 32 * q] = accM[q];
        pc[lane + 32 * q] = cntL[q]; pc[(256 + 32) + lane + 32 * q] = cntM[q];
    }
    pc[256 + lane] = kL; pc[(256 + 32) + 256 + lane] = kM;
}

// fixed-order sum over the warps' partials
__global__ void k5_reduce_kernel(const double* __restrict__ part_rate, const long long* __restrict__ part_cnt, int n_warps, int nb, int bin0,
                                 double* __restrict__ sum_rate, long long* __restrict__ shift_cnt, long long* __restrict__ k_cnt) {
    const int t = threadIdx.x;            // 0..255 bin of this pass (+ 32 threads' worth of K counts handled by t < 32)
    const int side = blockIdx.x;
    const int j = bin0 + t;
    double s = 0.0; long long c = 0, kc = 0;
    for (int w = 0; w < n_warps; ++w) {
        s += part_rate[((size_t)w * 2 + side) * 256 + t];
        c += part_cnt[((size_t)w * 2 + side) * (256 + 32) + t];
        if (t < 32) kc += part_cnt[((size_t)w * 2 + side) * (256 + 32) + 256 + t];
    }
    if (j < nb) { sum_rate[(size_t)side * nb + j] = s; shift_cnt[(size_t)side * nb + j] = c; }
    if (bin0 == 0 && t < 32) k_cnt[side * 32 + t] = kc;
}

}  // namespace

extern "C" int lr_summarize_records(lr_handle_t h, const double* d_records, int64_t n_records, double first_edge, int32_t n_bins,
                                    double* d_sum_rate, int64_t* d_shift_count, int64_t* d_k_count, void* stream) {
    LR_REQUIRE(h && d_sum_rate && d_shift_count && d_k_count && (n_records == 0 || d_records), "lr_summarize_records: null pointer");
    LR_REQUIRE(n_records >= 0 && n_bins >= 1, "lr_summarize_records: bad sizes");
    LR_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    const int ctas = h->sm_count * 2;
    const int n_warps = ctas * K5_WARPS_PER_CTA;
    const size_t rate_bytes = (size_t)n_warps * 2 * 256 * sizeof(double);
    const size_t cnt_bytes = (size_t)n_warps * 2 * (256 + 32) * sizeof(long long);
    int rc = lr_ws_reserve(h, rate_bytes + cnt_bytes);
    if (rc != LR_OK) return rc;
    K5Params p;
    p.rec = d_records; p.n_rec = n_records; p.first_edge = first_edge; p.nb = n_bins;
    p.part_rate = (double*)h->ws; p.part_cnt = (long long*)((char*)h->ws + rate_bytes);
    for (int bin0 = 0; bin0 < n_bins; bin0 += 256) {            // 256 bins per pass over the records
        p.bin0 = bin0;
        k5_accumulate_kernel<<<ctas, K5_WARPS_PER_CTA * 32, 0, st>>>(p);
        LR_CUDA(cudaGetLastError());
        k5_reduce_kernel<<<2, 256, 0, st>>>(p.part_rate, p.part_cnt, n_warps, n_bins, bin0, d_sum_rate, (long long*)d_shift_count, (long long*)d_k_count);
        LR_CUDA(cudaGetLastError());
        h->launches += 2;
    }
    return LR_OK;
}
