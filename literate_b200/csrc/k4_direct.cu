// K4: the Keiding birth-death log-likelihood evaluated DIRECTLY over the lineages, without binning -- the formulation
// of the reference's ancestor (other/LiteRateBDI_ext.py:124-160, get_BDlik / BD_partial_lik per lineage) and of the
// north star's "one block per (state, lineage tile)" sketch.  It is a VALIDATION path: the production path bins once (K1)
// and evaluates states in O(K) from prefix tables (K2/K3); this kernel re-reads the lineages for every group of states
// (16 B per lineage per group) and must give the same number,
//     sum_j [ sp_j log lam_j - lam_j br_j + ex_j log mu_j - mu_j br_j ]          (BD_lik_Keiding, LiteRateForward.py:137-148)
//   = sum_i { [ts_i in window] log lam(bin of ts_i) + [te_i in window] log mu(bin of te_i)
//             - integral over the lineage's time at risk of (lam + mu) },
// which is an independent check of the identity K1 rests on at full size.
//
// Shape: grid = (lineage tiles, state groups).  A CTA stages log lam, log mu and the prefix sums of (lam + mu) of its
// K4_GROUP states in shared memory, streams its tile of (ts, te) once with 128-bit loads and keeps K4_GROUP partial sums
// per thread in registers; block partials go to [tile][state] and a second kernel adds them in a fixed order, so the
// result is deterministic.
#include "lr_common.cuh"

namespace {

constexpr int K4_GROUP = 8;        // states per CTA
constexpr int K4_THREADS = 256;

struct K4Params {
    const double* ts;
    const double* te;
    long long n;
    long long per_tile;            // lineages per CTA (multiple of 2)
    int fb;
    int nb;
    const double* lam;             // [n_states][nb]
    const double* mu;
    int n_states;
    double* partial;               // [n_tiles][n_states]
    int vec_ok;
};

__device__ __forceinline__ void k4_lineage(double ts, double te, int fb, int nb, const double* s_ll, const double* s_lm,
                                           const double* s_r, const double* s_R, int ng, double (&acc)[K4_GROUP]) {
    if (!(te > ts)) {
        // no time at risk: the events still count (:120-121)
        const double T0 = (double)fb, T1 = (double)fb + (double)nb;
        if (ts >= T0 && ts < T1) {
            const int a = __double2int_rd(ts) - fb;
#pragma unroll
            for (int g = 0; g < K4_GROUP; ++g) if (g < ng) acc[g] += s_ll[g * nb + a];
        }
        if (te > T0 && te <= T1) {
            const int b = __double2int_ru(te) - 1 - fb;
#pragma unroll
            for (int g = 0; g < K4_GROUP; ++g) if (g < ng) acc[g] += s_lm[g * nb + b];
        }
        return;
    }
    const double T0 = (double)fb, T1 = (double)fb + (double)nb;
    if (ts >= T1 || !(te > T0)) return;                       // entirely outside the window
    const bool born_in = ts >= T0, died_in = te <= T1;
    const double s = (born_in ? ts : T0) - T0, e = (died_in ? te : T1) - T0;     // clipped, in bin units from the window start
    const int a = __double2int_rd(s);                          // first bin with time at risk
    int b = __double2int_ru(e) - 1;                            // last bin with time at risk
    if (b < a) b = a;
    const double fa = (double)(a + 1) - s;                     // time at risk inside bin a if the lineage leaves it
    const double fbk = e - (double)b;                          // time at risk inside bin b
#pragma unroll
    for (int g = 0; g < K4_GROUP; ++g) {
        if (g >= ng) break;
        const double* r = s_r + g * nb;
        const double* R = s_R + g * (nb + 1);
        double v = (a == b) ? -(e - s) * r[a] : -(fa * r[a] + (R[b] - R[a + 1]) + fbk * r[b]);
        if (born_in) v += s_ll[g * nb + a];
        if (died_in) v += s_lm[g * nb + b];
        acc[g] += v;
    }
}

__global__ void __launch_bounds__(K4_THREADS) k4_direct_kernel(const K4Params p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nb = p.nb, tid = threadIdx.x;
    const int s0 = blockIdx.y * K4_GROUP;
    const int ng = min(K4_GROUP, p.n_states - s0);
    double* s_ll = (double*)smem_raw;                  // [G][nb]   log lam
    double* s_lm = s_ll + K4_GROUP * nb;               // [G][nb]   log mu
    double* s_r = s_lm + K4_GROUP * nb;                // [G][nb]   lam + mu
    double* s_R = s_r + K4_GROUP * nb;                 // [G][nb+1] prefix sums of lam + mu
    for (int i = tid; i < ng * nb; i += blockDim.x) {
        const int g = i / nb, j = i - g * nb;
        const double l = p.lam[(size_t)(s0 + g) * nb + j], m = p.mu[(size_t)(s0 + g) * nb + j];
        s_ll[g * nb + j] = log(l); s_lm[g * nb + j] = log(m); s_r[g * nb + j] = l + m;
    }
    __syncthreads();
    if (tid < ng) {                                    // serial prefix per state: n_bins is a few hundred
        double a = 0.0;
        double* R = s_R + tid * (nb + 1);
        R[0] = 0.0;
        for (int j = 0; j < nb; ++j) { a += s_r[tid * nb + j]; R[j + 1] = a; }
    }
    __syncthreads();

    double acc[K4_GROUP];
#pragma unroll
    for (int g = 0; g < K4_GROUP; ++g) acc[g] = 0.0;
    const long long i0 = (long long)blockIdx.x * p.per_tile;
    long long i1 = i0 + p.per_tile;
    if (i1 > p.n) i1 = p.n;
    if (p.vec_ok) {
        const long long npair = (i1 - i0) / 2;
        const double2* t2 = (const double2*)(p.ts + i0);
        const double2* e2 = (const double2*)(p.te + i0);
        for (long long k = tid; k < npair; k += blockDim.x) {
            const double2 a = ld_stream_f64x2(t2 + k), b = ld_stream_f64x2(e2 + k);
            k4_lineage(a.x, b.x, p.fb, nb, s_ll, s_lm, s_r, s_R, ng, acc);
            k4_lineage(a.y, b.y, p.fb, nb, s_ll, s_lm, s_r, s_R, ng, acc);
        }
        if (((i1 - i0) & 1) && tid == 0) k4_lineage(p.ts[i1 - 1], p.te[i1 - 1], p.fb, nb, s_ll, s_lm, s_r, s_R, ng, acc);
    } else {
        for (long long i = i0 + tid; i < i1; i += blockDim.x)
            k4_lineage(ld_stream_f64(p.ts + i), ld_stream_f64(p.te + i), p.fb, nb, s_ll, s_lm, s_r, s_R, ng, acc);
    }
    // block reduction, fixed order
    __shared__ double red[K4_THREADS / 32][K4_GROUP];
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int g = 0; g < K4_GROUP; ++g) {
        const double v = warp_sum(acc[g]);
        if (lane == 0) red[warp][g] = v;
    }
    __syncthreads();
    if (tid < ng) {
        double v = 0.0;
        for (int w = 0; w < K4_THREADS / 32; ++w) v += red[w][tid];
        p.partial[(size_t)blockIdx.x * p.n_states + s0 + tid] = v;
    }
}

__global__ void k4_sum_kernel(const double* __restrict__ partial, int n_tiles, int n_states, double* __restrict__ out) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_states) return;
    double v = 0.0;
    for (int t = 0; t < n_tiles; ++t) v += partial[(size_t)t * n_states + s];
    out[s] = v;
}

}  // namespace

extern "C" int lr_loglik_direct(lr_handle_t h, const double* d_ts, const double* d_te, int64_t n, int64_t first_bin, int32_t n_bins,
                                const double* d_lam, const double* d_mu, int32_t n_states, double* d_out, void* stream) {
    LR_REQUIRE(h && d_lam && d_mu && d_out && (n == 0 || (d_ts && d_te)), "lr_loglik_direct: null pointer");
    LR_REQUIRE(n >= 0 && n_bins >= 1 && n_states >= 1, "lr_loglik_direct: bad sizes");
    const size_t smem = (size_t)K4_GROUP * (4 * (size_t)n_bins + 1) * sizeof(double);
    if (smem > (size_t)h->max_smem_optin - 2048 || first_bin <= -(1ll << 30) || first_bin >= (1ll << 30)) {
        lr_set_error("lr_loglik_direct: n_bins too large for the shared-memory tables (max %d) or first_bin out of range",
                     (int)(((size_t)h->max_smem_optin - 2048) / (K4_GROUP * 4 * sizeof(double))));
        return LR_ERR_UNSUPPORTED;
    }
    LR_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    const int groups = (n_states + K4_GROUP - 1) / K4_GROUP;
    int tiles = h->sm_count * 4 / (groups < 4 ? groups : 4);
    if (tiles < 1) tiles = 1;
    long long per_tile = (n + tiles - 1) / tiles;
    per_tile = (per_tile + 1) & ~1ll;
    if (per_tile < 2) per_tile = 2;
    tiles = n > 0 ? (int)((n + per_tile - 1) / per_tile) : 1;
    int rc = lr_ws_acquire(h, (size_t)tiles * n_states * sizeof(double), st);
    if (rc != LR_OK) return rc;
    K4Params p;
    p.ts = d_ts; p.te = d_te; p.n = n; p.per_tile = per_tile; p.fb = (int)first_bin; p.nb = n_bins;
    p.lam = d_lam; p.mu = d_mu; p.n_states = n_states; p.partial = (double*)h->ws;
    p.vec_ok = (((uintptr_t)d_ts | (uintptr_t)d_te) & 15) == 0;
    LR_CUDA(cudaFuncSetAttribute(k4_direct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k4_direct_kernel<<<dim3(tiles, groups), K4_THREADS, smem, st>>>(p);
    LR_CUDA(cudaGetLastError());
    k4_sum_kernel<<<(n_states + 127) / 128, 128, 0, st>>>((const double*)h->ws, tiles, n_states, d_out);
    LR_CUDA(cudaGetLastError());
    h->launches += 2;
    return LR_OK;
}
