// K5: posterior accumulators straight from device-resident sample records -- the per-bin marginal rates, shift-time
// histograms and number-of-rates counts that plotRJforward.v3.py derives from the text logs one row at a time
// (get_marginal_rates :92-139, get_r_plot :166-176, get_K_values :292-305).  SURVEY 8 f-1: for ensembles of thousands of
// chains the records never have to leave the GPU (nor be formatted as text) to be summarised.
//
//   marginal rate of bin j in one sample = rates[ number of shift times falling in bins 0..j ]      (np.histogram semantics:
//   half-open unit bins from `first_edge`, the last one closed, times outside ignored)
//
// One warp per record (grid-stride, fixed grid => fixed summation order => deterministic): lane k holds rate k and shift k,
// lane l owns the 8 consecutive bins 8 l .. 8 l + 7 of a pass of 256; partial sums stay in registers across records and go
// to [warp][...] partials that a second kernel adds in order.  Algorithmic traffic: 1152 B per record.
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

// Rate index of a bin = number of (valid) shifts at or before it.  Instead of comparing every bin with every shift
// (O(K) shuffles and 2 x 8 integer operations per shift and lane: 0.7 TB/s, ALU-bound) the shifts are SCATTERED: every slot
// that holds a valid shift adds 1 to a per-warp shared-memory mark of its bin, and an integer prefix sum over the bins
// (8 consecutive bins per lane + one warp scan) turns the marks into the index; the marks themselves are the histogram of
// the shift times.  The rates are read back from 32 shared doubles.  Integer arithmetic only up to the final lookup: exact.
// (Measured alternatives on 2.05 M records: compare-all 3.4 ms; this scatter 2.1 ms, 1.33 ms with the record loads made
// independent of K and issued ahead, 0.80 ms with the parallel reduce below; a 5-step binary search over the sorted shift
// bins by shuffles, no shared memory: 1.74 ms -- the 61 indexed shuffles per side cost more than the three barriers.)
struct K5Warp {
    int* mark;        // [256] shifts per bin of this pass
    double* rate;     // [32]
};

// what a lane needs of one record: six independent loads (no load waits for K), issued one record ahead of their use
struct K5Rec { double kl, km, rl, tl, rm, tm; };
__device__ __forceinline__ K5Rec k5_load(const double* r, int lane) {
    K5Rec x;
    x.kl = __ldg(r + REC_KL); x.km = __ldg(r + REC_KM);
    x.rl = __ldg(r + REC_L + lane); x.tl = __ldg(r + REC_TL + lane);
    x.rm = __ldg(r + REC_M + lane); x.tm = __ldg(r + REC_TM + lane);
    return x;
}

// One side of one record: v[q] = marginal rate and m[q] = number of shift times of the lane's bin bin0 + 8 lane + q.
__device__ __forceinline__ void k5_eval(double k_raw, double rate_raw, double t_raw, double e0, int nb, int bin0, int lane,
                                        const K5Warp& w, double (&v)[K5_BINS_PER_LANE], int (&m)[K5_BINS_PER_LANE], int& K) {
    K = (int)k_raw;
    const double rate = lane < K ? rate_raw : 0.0;
    if (K == 1) {                                   // no shift: one rate everywhere
        const double r0 = __shfl_sync(0xffffffffu, rate, 0);
#pragma unroll
        for (int q = 0; q < K5_BINS_PER_LANE; ++q) { v[q] = r0; m[q] = 0; }
        return;
    }
    // slot k >= 1 holds shift k-1; its histogram bin (np.histogram: half-open unit bins, the last one closed, outside ignored)
    int sb = 0x7fffffff;
    if (lane >= 1 && lane < K) {
        const double s = t_raw;
        if (s >= e0 && s <= e0 + (double)nb) { sb = __double2int_rd(s - e0); if (sb > nb - 1) sb = nb - 1; }
    }
    __syncwarp();                                   // the previous use of the marks has been read by every lane
    *reinterpret_cast<int4*>(w.mark + K5_BINS_PER_LANE * lane) = make_int4(0, 0, 0, 0);
    *reinterpret_cast<int4*>(w.mark + K5_BINS_PER_LANE * lane + 4) = make_int4(0, 0, 0, 0);
    w.rate[lane] = rate;
    __syncwarp();
    const int before = __popc(__ballot_sync(0xffffffffu, sb < bin0));            // shifts in earlier passes' bins
    if (sb >= bin0 && sb < bin0 + 256) atomicAdd(w.mark + (sb - bin0), 1);
    __syncwarp();
    const int4 m0 = *reinterpret_cast<const int4*>(w.mark + K5_BINS_PER_LANE * lane);
    const int4 m1 = *reinterpret_cast<const int4*>(w.mark + K5_BINS_PER_LANE * lane + 4);
    m[0] = m0.x; m[1] = m0.y; m[2] = m0.z; m[3] = m0.w; m[4] = m1.x; m[5] = m1.y; m[6] = m1.z; m[7] = m1.w;
    int tot = 0;
#pragma unroll
    for (int q = 0; q < K5_BINS_PER_LANE; ++q) tot += m[q];
    int run = tot;                                  // inclusive scan of the lanes' totals
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, run, o);
        if (lane >= o) run += t;
    }
    run += before - tot;                            // shifts before this lane's first bin
#pragma unroll
    for (int q = 0; q < K5_BINS_PER_LANE; ++q) {
        run += m[q];
        v[q] = w.rate[run & 31];
    }
}

// The accumulating twin of k5_eval (kept fused: going through v[] / m[] costs the accumulate kernel 10 % at its 128 registers).
__device__ __forceinline__ void k5_side(double k_raw, double rate_raw, double t_raw, double e0, int nb, int bin0, int lane,
                                        const K5Warp& w, double (&acc)[K5_BINS_PER_LANE], int (&cnt)[K5_BINS_PER_LANE], int& kcnt) {
    const int K = (int)k_raw;
    const double rate = lane < K ? rate_raw : 0.0;
    if (bin0 == 0 && lane == K - 1) kcnt++;
    if (K == 1) {                                   // no shift: one rate everywhere
        const double r0 = __shfl_sync(0xffffffffu, rate, 0);
#pragma unroll
        for (int q = 0; q < K5_BINS_PER_LANE; ++q)
            if (bin0 + K5_BINS_PER_LANE * lane + q < nb) acc[q] += r0;
        return;
    }
    int sb = 0x7fffffff;
    if (lane >= 1 && lane < K) {
        const double s = t_raw;
        if (s >= e0 && s <= e0 + (double)nb) { sb = __double2int_rd(s - e0); if (sb > nb - 1) sb = nb - 1; }
    }
    __syncwarp();
    *reinterpret_cast<int4*>(w.mark + K5_BINS_PER_LANE * lane) = make_int4(0, 0, 0, 0);
    *reinterpret_cast<int4*>(w.mark + K5_BINS_PER_LANE * lane + 4) = make_int4(0, 0, 0, 0);
    w.rate[lane] = rate;
    __syncwarp();
    const int before = __popc(__ballot_sync(0xffffffffu, sb < bin0));
    if (sb >= bin0 && sb < bin0 + 256) atomicAdd(w.mark + (sb - bin0), 1);
    __syncwarp();
    const int4 m0 = *reinterpret_cast<const int4*>(w.mark + K5_BINS_PER_LANE * lane);
    const int4 m1 = *reinterpret_cast<const int4*>(w.mark + K5_BINS_PER_LANE * lane + 4);
    const int m[K5_BINS_PER_LANE] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
    int tot = 0;
#pragma unroll
    for (int q = 0; q < K5_BINS_PER_LANE; ++q) tot += m[q];
    int run = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, run, o);
        if (lane >= o) run += t;
    }
    run += before - tot;
#pragma unroll
    for (int q = 0; q < K5_BINS_PER_LANE; ++q) {
        run += m[q];
        if (bin0 + K5_BINS_PER_LANE * lane + q < nb) { acc[q] += w.rate[run & 31]; cnt[q] += m[q]; }
    }
}

// The per-sample matrix behind the HPD intervals (get_marginal_rates, plotRJforward.v3.py:92-139): row = record, column = bin of
// [bin_lo, bin_lo + bin_cnt), record-major so that a warp writes its record's row in one piece.
struct K5Expand {
    const double* rec;
    long long n_rec;
    double first_edge;
    int nb, bin_lo, bin_cnt;
    double* birth;              // [n_rec][bin_cnt]
    double* death;
};

__global__ void __launch_bounds__(K5_WARPS_PER_CTA * 32) k5_expand_kernel(const K5Expand p) {
    __shared__ __align__(16) int mark_s[K5_WARPS_PER_CTA][256];
    __shared__ double rate_s[K5_WARPS_PER_CTA][32];
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    K5Warp w;
    w.mark = mark_s[threadIdx.x >> 5]; w.rate = rate_s[threadIdx.x >> 5];
    for (long long i = warp; i < p.n_rec; i += n_warps) {
        const K5Rec cur = k5_load(p.rec + (size_t)i * REC_W, lane);
        for (int bin0 = p.bin_lo; bin0 < p.bin_lo + p.bin_cnt; bin0 += 256) {
            double v[K5_BINS_PER_LANE];
            int m[K5_BINS_PER_LANE], K;
#pragma unroll
            for (int side = 0; side < 2; ++side) {
                k5_eval(side ? cur.km : cur.kl, side ? cur.rm : cur.rl, side ? cur.tm : cur.tl, p.first_edge, p.nb, bin0, lane, w, v, m, K);
                double* row = (side ? p.death : p.birth) + (size_t)i * p.bin_cnt;
#pragma unroll
                for (int q = 0; q < K5_BINS_PER_LANE; ++q) {
                    const int j = bin0 + K5_BINS_PER_LANE * lane + q;
                    if (j < p.bin_lo + p.bin_cnt && j < p.nb) row[j - p.bin_lo] = v[q];
                }
            }
        }
    }
}

__global__ void __launch_bounds__(K5_WARPS_PER_CTA * 32, 2) k5_accumulate_kernel(const K5Params p) {
    __shared__ __align__(16) int mark_s[K5_WARPS_PER_CTA][256];
    __shared__ double rate_s[K5_WARPS_PER_CTA][32];
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    K5Warp w;
    w.mark = mark_s[threadIdx.x >> 5]; w.rate = rate_s[threadIdx.x >> 5];
    double accL[K5_BINS_PER_LANE], accM[K5_BINS_PER_LANE];
    int cntL[K5_BINS_PER_LANE], cntM[K5_BINS_PER_LANE];
    int kL = 0, kM = 0;
#pragma unroll
    for (int q = 0; q < K5_BINS_PER_LANE; ++q) { accL[q] = 0.0; accM[q] = 0.0; cntL[q] = 0; cntM[q] = 0; }
    // two records in flight per warp beside the one being processed
    K5Rec n1, n2;
    if (warp < p.n_rec) n1 = k5_load(p.rec + (size_t)warp * REC_W, lane);
    if (warp + n_warps < p.n_rec) n2 = k5_load(p.rec + (size_t)(warp + n_warps) * REC_W, lane);
    for (long long i = warp; i < p.n_rec; i += n_warps) {
        const K5Rec cur = n1;
        n1 = n2;
        if (i + 2 * n_warps < p.n_rec) n2 = k5_load(p.rec + (size_t)(i + 2 * n_warps) * REC_W, lane);
        k5_side(cur.kl, cur.rl, cur.tl, p.first_edge, p.nb, p.bin0, lane, w, accL, cntL, kL);
        k5_side(cur.km, cur.rm, cur.tm, p.first_edge, p.nb, p.bin0, lane, w, accM, cntM, kM);
    }
    // lane owns the 8 consecutive bins 8 lane .. 8 lane + 7 of this pass
    double* pr = p.part_rate + (size_t)warp * 2 * 256;
    long long* pc = p.part_cnt + (size_t)warp * 2 * (256 + 32);
#pragma unroll
    for (int q = 0; q < K5_BINS_PER_LANE; ++q) {
        pr[K5_BINS_PER_LANE * lane + q] = accL[q]; pr[256 + K5_BINS_PER_LANE * lane + q] = accM[q];
        pc[K5_BINS_PER_LANE * lane + q] = cntL[q]; pc[(256 + 32) + K5_BINS_PER_LANE * lane + q] = cntM[q];
    }
    pc[256 + lane] = kL; pc[(256 + 32) + 256 + lane] = kM;
}

// Fixed-order sum over the warps' partials.  grid (side, group of 32 bins) x block (32 bins, 32 slices): slice g adds the
// partials of warps g, g + 32, ... in order, then one thread per bin adds the 32 slice sums in order -- deterministic, and the
// serial part is n_warps / 32 loads deep instead of n_warps.
__global__ void __launch_bounds__(1024) k5_reduce_kernel(const double* __restrict__ part_rate, const long long* __restrict__ part_cnt,
                                                          int n_warps, int nb, int bin0, double* __restrict__ sum_rate,
                                                          long long* __restrict__ shift_cnt, long long* __restrict__ k_cnt) {
    __shared__ double s_rate[32][33];
    __shared__ long long s_cnt[32][33], s_k[32][33];
    const int b = threadIdx.x, g = threadIdx.y;
    const int side = blockIdx.x, t = blockIdx.y * 32 + b;         // t: bin of this pass
    const bool with_k = blockIdx.y == 0;                           // the first group also adds the K counts (32 of them)
    double s = 0.0; long long c = 0, kc = 0;
#pragma unroll 4
    for (int w = g; w < n_warps; w += 32) {
        s += part_rate[((size_t)w * 2 + side) * 256 + t];
        c += part_cnt[((size_t)w * 2 + side) * (256 + 32) + t];
        if (with_k) kc += part_cnt[((size_t)w * 2 + side) * (256 + 32) + 256 + b];
    }
    s_rate[g][b] = s; s_cnt[g][b] = c; s_k[g][b] = kc;
    __syncthreads();
    if (g == 0) {
        double S = 0.0; long long C = 0, KC = 0;
        for (int i = 0; i < 32; ++i) { S += s_rate[i][b]; C += s_cnt[i][b]; KC += s_k[i][b]; }
        const int j = bin0 + t;
        if (j < nb) { sum_rate[(size_t)side * nb + j] = S; shift_cnt[(size_t)side * nb + j] = C; }
        if (bin0 == 0 && with_k) k_cnt[side * 32 + b] = KC;
    }
}

}  // namespace

extern "C" int lr_marginal_rates(lr_handle_t h, const double* d_records, int64_t n_records, double first_edge, int32_t n_bins,
                                 int32_t bin_lo, int32_t bin_cnt, double* d_birth, double* d_death, void* stream) {
    LR_REQUIRE(h && d_birth && d_death && (n_records == 0 || d_records), "lr_marginal_rates: null pointer");
    LR_REQUIRE(n_records >= 0 && n_bins >= 1 && bin_lo >= 0 && bin_cnt >= 1 && bin_lo + bin_cnt <= n_bins, "lr_marginal_rates: bad sizes");
    if (n_records == 0) return LR_OK;
    LR_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    K5Expand p;
    p.rec = d_records; p.n_rec = n_records; p.first_edge = first_edge; p.nb = n_bins; p.bin_lo = bin_lo; p.bin_cnt = bin_cnt;
    p.birth = d_birth; p.death = d_death;
    k5_expand_kernel<<<h->sm_count * 4, K5_WARPS_PER_CTA * 32, 0, st>>>(p);
    LR_CUDA(cudaGetLastError());
    h->launches += 1;
    return LR_OK;
}

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
    int rc = lr_ws_acquire(h, rate_bytes + cnt_bytes, st);
    if (rc != LR_OK) return rc;
    K5Params p;
    p.rec = d_records; p.n_rec = n_records; p.first_edge = first_edge; p.nb = n_bins;
    p.part_rate = (double*)h->ws; p.part_cnt = (long long*)((char*)h->ws + rate_bytes);
    for (int bin0 = 0; bin0 < n_bins; bin0 += 256) {            // 256 bins per pass over the records
        p.bin0 = bin0;
        k5_accumulate_kernel<<<ctas, K5_WARPS_PER_CTA * 32, 0, st>>>(p);
        LR_CUDA(cudaGetLastError());
        k5_reduce_kernel<<<dim3(2, 8), dim3(32, 32), 0, st>>>(p.part_rate, p.part_cnt, n_warps, n_bins, bin0, d_sum_rate, (long long*)d_shift_count, (long long*)d_k_count);
        LR_CUDA(cudaGetLastError());
        h->launches += 2;
    }
    return LR_OK;
}


// ------------------------------------------------------------------------------------------------
// Envelopes over imputation replicates (utilities/imputation_averager.py:23-61): per bin, mean / min / max over the
// replicates of the empirical rates sp/br, ex/br and of br itself, straight from K1's device output.  One thread per
// bin walks the replicates in order -- the accumulation order of numpy's mean(axis=0) over the stacked div.log tables,
// so the means carry numpy's bits.  out: [9][n_bins] = death mean/min/max, birth mean/min/max, diversity mean/min/max.
// ------------------------------------------------------------------------------------------------
namespace {
__global__ void k5_envelope_kernel(const long long* __restrict__ sp, const long long* __restrict__ ex, const double* __restrict__ br,
                                   int n_rep, int nb, double* __restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nb) return;
    double s[3] = {0.0, 0.0, 0.0}, lo[3], hi[3];
    for (int r = 0; r < n_rep; ++r) {
        const double k = br[(size_t)r * nb + j];
        const double v[3] = {(double)ex[(size_t)r * nb + j] / k, (double)sp[(size_t)r * nb + j] / k, k};
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            s[q] += v[q];
            // numpy's minimum / maximum propagate NaN (0/0 in an empty bin)
            if (r == 0) { lo[q] = v[q]; hi[q] = v[q]; }
            else {
                lo[q] = (v[q] != v[q] || lo[q] != lo[q]) ? (v[q] != v[q] ? v[q] : lo[q]) : (v[q] < lo[q] ? v[q] : lo[q]);
                hi[q] = (v[q] != v[q] || hi[q] != hi[q]) ? (v[q] != v[q] ? v[q] : hi[q]) : (v[q] > hi[q] ? v[q] : hi[q]);
            }
        }
    }
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        out[(size_t)(3 * q + 0) * nb + j] = s[q] / (double)n_rep;
        out[(size_t)(3 * q + 1) * nb + j] = lo[q];
        out[(size_t)(3 * q + 2) * nb + j] = hi[q];
    }
}
}  // namespace

extern "C" int lr_imputation_envelope(lr_handle_t h, const int64_t* d_sp, const int64_t* d_ex, const double* d_br, int32_t n_rep, int32_t n_bins,
                                      double* d_out, void* stream) {
    LR_REQUIRE(h && d_sp && d_ex && d_br && d_out, "lr_imputation_envelope: null pointer");
    LR_REQUIRE(n_rep >= 1 && n_bins >= 1, "lr_imputation_envelope: bad sizes");
    LR_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    k5_envelope_kernel<<<(n_bins + 127) / 128, 128, 0, st>>>((const long long*)d_sp, (const long long*)d_ex, d_br, n_rep, n_bins, d_out);
    LR_CUDA(cudaGetLastError());
    h->launches += 1;
    return LR_OK;
}
