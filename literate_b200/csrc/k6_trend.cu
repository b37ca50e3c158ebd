// K6: TrendRate -- fixed-dimension Metropolis-Hastings chains on the binned statistics of K1 with
// birth and death rates that follow an exogenous trend.  Replaces the loop of trend_rate.py:102-196 and the
// literate_library.py functions it calls (proposals :140-165, priors :182-187, adequacy :268-279).
//
//   lambda_j = l_min + alpha * trend_j ** delta      mu_j = m_min + beta * trend_j ** gamma        (trend_rate.py:73-91)
//   log-lik  = sum_j log(lambda_j) sp_j - lambda_j br_j  +  sum_j log(mu_j) ex_j - mu_j br_j       (Keiding form, :82, :88)
//
// One warp = one chain, whole loop on device.  Lane p < 6 owns parameter p: it draws the parameter's Bernoulli mask and
// its multiplier / normal step from one Philox call keyed by (seed, chain, iteration, lane), evaluates the parameter's
// prior difference (no logarithm: the Gamma(3) parameters only move by multipliers), and the six values are broadcast by
// shuffles; lane 6 draws the branch uniform and the acceptance uniform (accept test through a single-precision bracket of log u).
// Bins are strided over the 32 lanes (bin j lives in lane j % 32; the first 32 bins' statistics stay in registers), the
// two side sums come out of one butterfly.  A side none of whose parameters was touched keeps its stored likelihood (the
// reference recomputes the identical number).  Nothing but the read-only per-bin table (L1-resident) is read inside the
// loop; the only global writes are the sample records.
#include <stdlib.h>
#include "lr_common.cuh"

#define TR_NPAR 6
#define TR_ROWS 6          // per replicate: sp, ex, br, ln(trend), empirical birth rate, empirical death rate
#define TR_SP 0
#define TR_EX 1
#define TR_BR 2
#define TR_LNT 3
#define TR_XB 4
#define TR_XD 5
#define TR_SMALL 0.000000000000001      // SMALL_NUMBER, trend_rate.py:55
#define TR_LN_MULT 0.19062035960864987  // 2*log(1.1): update_multiplier_proposal_vec(d=1.1), literate_library.py:160
#define TR_NORM_SD 0.001                // update_normal_nobound_vec(d=.001), trend_rate.py:168

struct TrendChain {
    double p[TR_NPAR];     // l_min, m_min, alpha, beta, delta, gamma (trend_rate.py:74)
    double likB, likD, prior;
    long long it;          // iterations done
    long long accepted;
    unsigned chain;        // global chain id (Philox key)
    int rep;
};

struct lr_trend_s {
    lr_handle_t h;
    lr_last_stream last;           // stream that touched the object last (lr_order)
    int n_rep, n_bins, nbp;        // nbp: row pitch of the table
    int const_b, const_d;
    double* tab;                   // device [n_rep][TR_ROWS][nbp]
    double* cst;                   // device [n_rep][2]: sum x, sum x^2 of the empirical rates (adequacy)
    int n_chains;
    uint64_t seed;
    TrendChain* st;                // device [n_chains]
    double f_mult[TR_NPAR], f_norm[TR_NPAR];
};

namespace {

struct TrendView {
    const double* tab;
    int nb, nbp, const_b, const_d;
    double Sx, Sxx;
    // the bins of the likelihood sums this warp owns: first + lane (statistics in registers), then every `stride`-th after it.
    // One warp per chain: first = 0, stride = 32.  W warps per chain (wide build): first = 32 w, stride = 32 W.
    int first, stride;
    double sp0, ex0, br0, lnT0;
};
// warps per chain of the wide build, by the number of bins
__host__ __device__ inline int trend_groups(int nb) { return nb > 128 ? 4 : (nb > 64 ? 2 : 1); }

__device__ __forceinline__ TrendView trend_view(const double* tab_all, const double* cst_all, int rep, int nb, int nbp,
                                                int const_b, int const_d, int lane, int first = 0, int stride = 32) {
    TrendView v;
    v.tab = tab_all + (size_t)rep * TR_ROWS * nbp;
    v.nb = nb; v.nbp = nbp; v.const_b = const_b; v.const_d = const_d;
    v.Sx = cst_all[2 * rep]; v.Sxx = cst_all[2 * rep + 1];
    v.first = first; v.stride = stride;
    const int j = first + lane;
    const bool on = j < nb;
    v.sp0 = on ? v.tab[TR_SP * nbp + j] : 0.0;
    v.ex0 = on ? v.tab[TR_EX * nbp + j] : 0.0;
    v.br0 = on ? v.tab[TR_BR * nbp + j] : 0.0;
    v.lnT0 = on ? v.tab[TR_LNT * nbp + j] : 0.0;
    return v;
}

// rate of one bin (trend_rate.py:75-80 / :83-87): floor = a + b * T**c, values <= 0 become SMALL_NUMBER
__device__ __forceinline__ double trend_rate(double a, double b, double c, double lnT, int is_const) {
    if (is_const) return a;
    double r = a + b * exp(c * lnT);
    return r > 0.0 ? r : TR_SMALL;
}

// T**delta and T**gamma of the first 32 bins (one per lane): they change only when delta / gamma are proposed
struct PowCache { double b, d; };

// both side likelihoods (uniform over the warp); a side with do_* == false keeps the value passed in.  newPowB/newPowD: the
// exponent of that side differs from the one `pc` was computed for (pc is updated in place)
// `extra`: a per-lane term (Hastings share + prior difference of the lane's parameter) that rides the same butterfly; its
// warp total comes back in place.
// partial sums over the bins this warp owns (per lane, not yet reduced)
// termB / termD (wide build): if not null, every bin's term is also stored at its bin index
__device__ __forceinline__ void trend_lik_partial(const TrendView& v, const double* p, int lane, bool doB, bool doD, bool newPowB, bool newPowD,
                                                  PowCache& pc, double& sB, double& sD, double* termB = nullptr, double* termD = nullptr) {
    sB = 0.0; sD = 0.0;
    if (doB) {
        double lam = p[0];
        if (!v.const_b) {
            if (newPowB) pc.b = exp(p[4] * v.lnT0);
            lam = p[0] + p[2] * pc.b;
            lam = lam > 0.0 ? lam : TR_SMALL;
        }
        sB = log(lam) * v.sp0 - lam * v.br0;
    }
    if (doD) {
        double mu = p[1];
        if (!v.const_d) {
            if (newPowD) pc.d = exp(p[5] * v.lnT0);
            mu = p[1] + p[3] * pc.d;
            mu = mu > 0.0 ? mu : TR_SMALL;
        }
        sD = log(mu) * v.ex0 - mu * v.br0;
    }
    if (termB != nullptr && v.first + lane < v.nb) { termB[v.first + lane] = sB; termD[v.first + lane] = sD; }
    for (int j = v.first + lane + v.stride; j < v.nb; j += v.stride) {
        const double lnT = __ldg(v.tab + TR_LNT * v.nbp + j), br = __ldg(v.tab + TR_BR * v.nbp + j);
        double tB = 0.0, tD = 0.0;
        if (doB) {
            const double lam = trend_rate(p[0], p[2], p[4], lnT, v.const_b);
            tB = log(lam) * __ldg(v.tab + TR_SP * v.nbp + j) - lam * br;
            sB += tB;
        }
        if (doD) {
            const double mu = trend_rate(p[1], p[3], p[5], lnT, v.const_d);
            tD = log(mu) * __ldg(v.tab + TR_EX * v.nbp + j) - mu * br;
            sD += tD;
        }
        if (termB != nullptr) { termB[j] = tB; termD[j] = tD; }
    }
}

__device__ __forceinline__ void trend_lik(const TrendView& v, const double* p, int lane, bool doB, bool doD, bool newPowB, bool newPowD,
                                          PowCache& pc, double& likB, double& likD, double& extra) {
    double sB, sD;
    trend_lik_partial(v, p, lane, doB, doD, newPowB, newPowD, pc, sB, sD);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sB += __shfl_xor_sync(0xffffffffu, sB, o);
        sD += __shfl_xor_sync(0xffffffffu, sD, o);
        extra += __shfl_xor_sync(0xffffffffu, extra, o);
    }
    if (doB) likB = sB;
    if (doD) likD = sD;
}

// prior term of parameter `lane` (calc_prior, trend_rate.py:93-100; closed forms of scipy's gamma/norm logpdf)
__device__ __forceinline__ double trend_prior_term(int lane, double x) {
    if (lane < 2) {                         // Gamma(a=1, scale 10, loc .001)
        const double y = x - 0.001;
        return y < 0.0 ? -INFINITY : -2.302585092994046 - y / 10.0;
    }
    if (lane < 4) return -(x * x) / 50.0 - 2.528376445638773;      // Normal(0, 5): -log(5) - log(2 pi)/2
    if (lane < 6) return x > 0.0 ? 2.0 * log(2.0 * x) - 2.0 * x : -INFINITY;     // Gamma(a=3, scale .5): lgamma(3) = -log(.5)
    return 0.0;
}

// prior(x') - prior(x) of parameter `lane` inside the loop, without logarithms: the Gamma(3) parameters (delta, gamma) only move
// by multipliers, where log x' - log x is the log-multiplier `lm` itself.  `x` lies in the support (it is an accepted state).
__device__ __forceinline__ double trend_prior_delta(int lane, double x, double xn, double lm) {
    const double d = xn - x;
    if (lane < 2) return xn < 0.001 ? -INFINITY : -d / 10.0;
    if (lane < 4) return -(d * (xn + x)) / 50.0;
    if (lane < 6) return xn > 0.0 ? 2.0 * lm - 2.0 * d : -INFINITY;
    return 0.0;
}

// One proposal (trend_rate.py:166-171).  Lane p < 6 returns the proposed value of parameter p given the parameter's own
// value `x`, its mask bit and its draw (uniform for the multiplier, standard normal for the additive step); `hast` receives
// the lane's share of the Hastings ratio.
__device__ __forceinline__ double trend_propose(int kind_normal, double x, bool on, double draw, double& hast) {
    hast = 0.0;
    if (!on) return x;
    if (kind_normal) return x + TR_NORM_SD * draw;          // literate_library.py:140-146
    const double lm = TR_LN_MULT * (draw - 0.5);            // literate_library.py:156-165: m = exp(l (u - .5)), U = sum log m
    hast = lm;
    return x * exp_small(lm);
}

__device__ __forceinline__ void bcast6(double mine, double* p) {
#pragma unroll
    for (int k = 0; k < TR_NPAR; ++k) p[k] = __shfl_sync(0xffffffffu, mine, k);
}

// calculate_r_squared (literate_library.py:268-279) in closed form; x = empirical rates, y = the model's rates
__device__ __forceinline__ void trend_adequacy(const TrendView& v, const double* p, int lane, double out[3]) {
    double sy = 0, syy = 0, sxy = 0;
    for (int j = lane; j < v.nb; j += 32) {
        const double lnT = __ldg(v.tab + TR_LNT * v.nbp + j);
        const double lam = trend_rate(p[0], p[2], p[4], lnT, v.const_b), mu = trend_rate(p[1], p[3], p[5], lnT, v.const_d);
        const double xb = __ldg(v.tab + TR_XB * v.nbp + j), xd = __ldg(v.tab + TR_XD * v.nbp + j);
        sy += lam + mu; syy += lam * lam + mu * mu; sxy += lam * xb + mu * xd;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sy += __shfl_xor_sync(0xffffffffu, sy, o);
        syy += __shfl_xor_sync(0xffffffffu, syy, o);
        sxy += __shfl_xor_sync(0xffffffffu, sxy, o);
    }
    const double n = 2.0 * v.nb;
    const double c = sxy / v.Sxx;
    const double ssres = syy - c * sxy;
    const double var_f = c * c * (v.Sxx - v.Sx * v.Sx / n) / (n - 1.0);
    const double sres = sy - c * v.Sx;
    const double var_r = (ssres - sres * sres / n) / (n - 1.0);
    out[0] = c; out[1] = 1.0 - ssres / syy; out[2] = var_f / (var_f + var_r);
}

// record: [0] it [1] likelihood [2] likelihood_birth [3] likelihood_death [4] prior [5..10] parameters [11..13] adequacy
//         [14] accepted so far [15] reserved [16 .. 16+nb) birth rates [16+nb .. 16+2nb) death rates
__device__ __forceinline__ void trend_record(double* rec, const TrendView& v, const double* p, double mine, double likB, double likD,
                                             long long it, long long accepted, int lane) {
    double adq[3];
    trend_adequacy(v, p, lane, adq);
    const double prior = warp_sum(trend_prior_term(lane, mine));
    if (lane == 0) {
        rec[0] = (double)it; rec[1] = likB + likD; rec[2] = likB; rec[3] = likD; rec[4] = prior;
        rec[11] = adq[0]; rec[12] = adq[1]; rec[13] = adq[2]; rec[14] = (double)accepted; rec[15] = 0.0;
    }
#pragma unroll
    for (int k = 0; k < TR_NPAR; ++k)
        if (lane == k) rec[5 + k] = p[k];
    for (int j = lane; j < v.nb; j += 32) {
        const double lnT = __ldg(v.tab + TR_LNT * v.nbp + j);
        rec[LR_TREND_REC_HEAD + j] = trend_rate(p[0], p[2], p[4], lnT, v.const_b);
        rec[LR_TREND_REC_HEAD + v.nb + j] = trend_rate(p[1], p[3], p[5], lnT, v.const_d);
    }
}

struct TrendRun {
    TrendChain* st;
    int n_chains;
    const double* tab;
    const double* cst;
    int nb, nbp, const_b, const_d;
    uint32_t k0, k1;
    long long n_iter, sample_every;
    double* records;       // [sample][chain][rec_doubles] or null
    int rec_doubles;
    double fd_mult[TR_NPAR], fd_norm[TR_NPAR];  // Bernoulli probability of every parameter under the two proposal kinds
};

__global__ void __launch_bounds__(128, 4) k6_trend_kernel(const TrendRun P) {
    const int lane = threadIdx.x & 31;
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= P.n_chains) return;
    TrendChain* S = P.st + c;
    const TrendView v = trend_view(P.tab, P.cst, S->rep, P.nb, P.nbp, P.const_b, P.const_d, lane);
    const unsigned chain = S->chain;
    double mine = lane < TR_NPAR ? S->p[lane] : 0.0;       // this lane's parameter; p[] = the broadcast copies
    double p[TR_NPAR];
    bcast6(mine, p);
    double likB = S->likB, likD = S->likD;
    PowCache pc;
    pc.b = v.const_b ? 1.0 : exp(p[4] * v.lnT0);
    pc.d = v.const_d ? 1.0 : exp(p[5] * v.lnT0);
    long long it = S->it, accepted = S->accepted;
    const long long it_end = it + P.n_iter;
    // this lane's Bernoulli probabilities (lanes >= 6: never on)
    const double fm = lane < TR_NPAR ? P.fd_mult[lane] : 0.0, fn = lane < TR_NPAR ? P.fd_norm[lane] : 0.0;
    long long next_sample = (it + P.sample_every - 1) / P.sample_every * P.sample_every;
    long long rec_idx = 0;

    for (; it < it_end; ++it) {
        const Philox4 r = philox4x32_10((uint32_t)it, (uint32_t)((unsigned long long)it >> 32), (uint32_t)lane | (0x60u << 8), chain, P.k0, P.k1);
        // lane 6: branch and acceptance uniforms
        const double ua = u01(r.x, r.y), ub = u01(r.z, r.w);
        const double rr = __shfl_sync(0xffffffffu, ua, 6);
        const double u_acc = __shfl_sync(0xffffffffu, ub, 6);
        const int kind_normal = rr < 0.33;                                  // trend_rate.py:167
        // lanes 0..5: mask from 32 bits, draw from 52 bits (+ 24 bits for the angle of Box-Muller)
        const double um = ((double)r.x + 0.5) * 2.3283064365386963e-10;
        const bool on = um < (kind_normal ? fn : fm);
        double draw = u01(r.y, r.z);
        if (kind_normal) draw = normal_f32(draw, r.w);
        double h;
        const double prop = trend_propose(kind_normal, mine, on, draw, h);
        const unsigned touched = __ballot_sync(0xffffffffu, on);
        double q[TR_NPAR];
        bcast6(prop, q);
        // Hastings share and prior difference of this lane's parameter: summed in the butterfly of the likelihood
        double hp = h + trend_prior_delta(lane, mine, prop, h);
        double nB = likB, nD = likD;
        PowCache npc = pc;
        trend_lik(v, q, lane, (touched & 0x15u) != 0, (touched & 0x2au) != 0, (touched & 0x10u) != 0, (touched & 0x20u) != 0, npc, nB, nD, hp);
        const double x = ((nB + nD) - (likB + likD)) + hp;
        if (it == 0 || mh_accept_gt(x, u_acc)) {                            // trend_rate.py:176
#pragma unroll
            for (int k = 0; k < TR_NPAR; ++k) p[k] = q[k];
            mine = prop;
            likB = nB; likD = nD; pc = npc;
            ++accepted;
        }
        if (it == next_sample) {                                             // trend_rate.py:184
            if (P.records) trend_record(P.records + ((size_t)rec_idx * P.n_chains + c) * P.rec_doubles, v, p, mine, likB, likD, it, accepted, lane);
            ++rec_idx;
            next_sample += P.sample_every;
        }
    }
    const double prior = warp_sum(trend_prior_term(lane, mine));
    if (lane < TR_NPAR) S->p[lane] = mine;
    if (lane == 0) { S->likB = likB; S->likD = likD; S->prior = prior; S->it = it; S->accepted = accepted; }
}

// WIDE build: W warps (one CTA) per chain.  With a few hundred chains and hundreds of bins one warp per chain leaves the GPU
// idle (ncu, round 1: sm__throughput 10 %, 7 bins per lane in series, each an exp -> log chain).  Here warp w evaluates bins
// 32 w + lane (+ 32 W k) -- at 200 bins and W = 4 two per lane instead of seven.  Every warp draws the same random numbers, builds
// the same proposal and takes the same accept decision; the warps exchange their PER-BIN terms through shared memory
// (double-buffered by iteration parity: one __syncthreads per iteration) and every warp then adds them up exactly as the
// one-warp kernel does -- lane l its bins l, l + 32, l + 64, ... in ascending order, then one butterfly -- so the chain is
// the one-warp kernel's chain, bit for bit (tests/test_gpu_trend.py), and does not depend on scheduling.
template <int W>
__global__ void __launch_bounds__(W * 32, 2) k6_trend_wide_kernel(const TrendRun P) {
    extern __shared__ double terms[];                  // [2 parities][2 sides][nbp]
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int c = blockIdx.x;
    if (c >= P.n_chains) return;
    TrendChain* S = P.st + c;
    const TrendView v = trend_view(P.tab, P.cst, S->rep, P.nb, P.nbp, P.const_b, P.const_d, lane, 32 * w, 32 * W);
    const unsigned chain = S->chain;
    double mine = lane < TR_NPAR ? S->p[lane] : 0.0;
    double p[TR_NPAR];
    bcast6(mine, p);
    double likB = S->likB, likD = S->likD;
    PowCache pc;
    pc.b = v.const_b ? 1.0 : exp(p[4] * v.lnT0);
    pc.d = v.const_d ? 1.0 : exp(p[5] * v.lnT0);
    long long it = S->it, accepted = S->accepted;
    const long long it_end = it + P.n_iter;
    const double fm = lane < TR_NPAR ? P.fd_mult[lane] : 0.0, fn = lane < TR_NPAR ? P.fd_norm[lane] : 0.0;
    long long next_sample = (it + P.sample_every - 1) / P.sample_every * P.sample_every;
    long long rec_idx = 0;
    __syncthreads();                                   // every warp has read the chain's state before warp 0 may rewrite it

    for (; it < it_end; ++it) {
        const Philox4 r = philox4x32_10((uint32_t)it, (uint32_t)((unsigned long long)it >> 32), (uint32_t)lane | (0x60u << 8), chain, P.k0, P.k1);
        const double ua = u01(r.x, r.y), ub = u01(r.z, r.w);
        const double rr = __shfl_sync(0xffffffffu, ua, 6);
        const double u_acc = __shfl_sync(0xffffffffu, ub, 6);
        const int kind_normal = rr < 0.33;                                  // trend_rate.py:167
        const double um = ((double)r.x + 0.5) * 2.3283064365386963e-10;
        const bool on = um < (kind_normal ? fn : fm);
        double draw = u01(r.y, r.z);
        if (kind_normal) draw = normal_f32(draw, r.w);
        double h;
        const double prop = trend_propose(kind_normal, mine, on, draw, h);
        const unsigned touched = __ballot_sync(0xffffffffu, on);
        double q[TR_NPAR];
        bcast6(prop, q);
        double hp = h + trend_prior_delta(lane, mine, prop, h);
        const bool doB = (touched & 0x15u) != 0, doD = (touched & 0x2au) != 0;
        PowCache npc = pc;
        double* tB = terms + (size_t)(it & 1) * 2 * P.nbp;
        double* tD = tB + P.nbp;
        double sB, sD;
        trend_lik_partial(v, q, lane, doB, doD, (touched & 0x10u) != 0, (touched & 0x20u) != 0, npc, sB, sD, tB, tD);
        __syncthreads();
        // the one-warp kernel's order: lane l adds its bins l, l + 32, ... ascending (0.0 for a side that was not evaluated)
        sB = 0.0; sD = 0.0;
        for (int j = lane; j < P.nb; j += 32) { sB += tB[j]; sD += tD[j]; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sB += __shfl_xor_sync(0xffffffffu, sB, o);
            sD += __shfl_xor_sync(0xffffffffu, sD, o);
            hp += __shfl_xor_sync(0xffffffffu, hp, o);
        }
        const double nB = doB ? sB : likB, nD = doD ? sD : likD;
        const double x = ((nB + nD) - (likB + likD)) + hp;
        if (it == 0 || mh_accept_gt(x, u_acc)) {                            // trend_rate.py:176
#pragma unroll
            for (int k = 0; k < TR_NPAR; ++k) p[k] = q[k];
            mine = prop;
            likB = nB; likD = nD; pc = npc;
            ++accepted;
        }
        if (it == next_sample) {                                             // trend_rate.py:184
            if (P.records && w == 0) {
                const TrendView v0 = trend_view(P.tab, P.cst, S->rep, P.nb, P.nbp, P.const_b, P.const_d, lane);
                trend_record(P.records + ((size_t)rec_idx * P.n_chains + c) * P.rec_doubles, v0, p, mine, likB, likD, it, accepted, lane);
            }
            ++rec_idx;
            next_sample += P.sample_every;
        }
    }
    if (w == 0) {
        const double prior = warp_sum(trend_prior_term(lane, mine));
        if (lane < TR_NPAR) S->p[lane] = mine;
        if (lane == 0) { S->likB = likB; S->likD = likD; S->prior = prior; S->it = it; S->accepted = accepted; }
    }
}

// initial state (trend_rate.py:141-156)
__global__ void k6_init_kernel(TrendChain* st, int n_chains, const int* rep_of_chain, long long chain_id0, const double* tab,
                               const double* cst, int nb, int nbp, int const_b, int const_d) {
    const int lane = threadIdx.x & 31;
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= n_chains) return;
    const int rep = rep_of_chain ? rep_of_chain[c] : 0;
    const TrendView v = trend_view(tab, cst, rep, nb, nbp, const_b, const_d, lane);
    const double mine = lane < 2 ? 0.1 : (lane < 4 ? 0.0 : (lane < 6 ? 1.0 : 0.0));
    double p[TR_NPAR];
    bcast6(mine, p);
    double likB = 0, likD = 0;
    PowCache pc;
    double unused = 0.0;
    trend_lik(v, p, lane, true, true, true, true, pc, likB, likD, unused);
    double pr = trend_prior_term(lane, mine);
    pr = warp_sum(pr);
    if (lane < TR_NPAR) st[c].p[lane] = mine;
    if (lane == 0) {
        st[c].likB = likB; st[c].likD = likD; st[c].prior = pr; st[c].it = 0; st[c].accepted = 0;
        st[c].chain = (unsigned)(chain_id0 + c); st[c].rep = rep;
    }
}

// per-replicate table: statistics as doubles, ln(trend), empirical rates and their sums
__global__ void k6_build_tables(const long long* __restrict__ sp, const long long* __restrict__ ex, const double* __restrict__ br,
                                const double* __restrict__ trend, int nb, int nbp, double* __restrict__ tab_all, double* __restrict__ cst_all) {
    const int rep = blockIdx.x;
    double* tab = tab_all + (size_t)rep * TR_ROWS * nbp;
    for (int j = threadIdx.x; j < nbp; j += blockDim.x) {
        const bool in = j < nb;
        const double U = in ? (double)sp[(size_t)rep * nb + j] : 0.0, D = in ? (double)ex[(size_t)rep * nb + j] : 0.0;
        const double K = in ? br[(size_t)rep * nb + j] : 0.0;
        tab[TR_SP * nbp + j] = U; tab[TR_EX * nbp + j] = D; tab[TR_BR * nbp + j] = K;
        tab[TR_LNT * nbp + j] = in ? log(trend[j]) : 0.0;
        tab[TR_XB * nbp + j] = in ? U / K : 0.0;
        tab[TR_XD * nbp + j] = in ? D / K : 0.0;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double sx = 0, sxx = 0;
        for (int j = 0; j < nb; ++j) {
            const double xb = tab[TR_XB * nbp + j], xd = tab[TR_XD * nbp + j];
            sx += xb + xd; sxx += xb * xb + xd * xd;
        }
        cst_all[2 * rep] = sx; cst_all[2 * rep + 1] = sxx;
    }
}

// parity entry: one warp per explicit parameter vector, optionally after one proposal with explicit draws
__global__ void k6_eval_kernel(const double* tab, const double* cst, int nb, int nbp, int const_b, int const_d, int n,
                               const int* rep, const double* params, const int* kind, const int* on, const double* draw,
                               double* out_params, double* out_hast, double* lik, double* prior, double* rates, double* adequacy) {
    const int lane = threadIdx.x & 31;
    const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (s >= n) return;
    const TrendView v = trend_view(tab, cst, rep ? rep[s] : 0, nb, nbp, const_b, const_d, lane);
    double mine = lane < TR_NPAR ? params[s * TR_NPAR + lane] : 0.0, h = 0.0;
    if (kind) mine = trend_propose(kind[s], mine, lane < TR_NPAR && on[s * TR_NPAR + lane] != 0, lane < TR_NPAR ? draw[s * TR_NPAR + lane] : 0.0, h);
    h = warp_sum(h);
    double p[TR_NPAR];
    bcast6(mine, p);
    double likB = 0, likD = 0;
    PowCache pc;
    double unused = 0.0;
    trend_lik(v, p, lane, true, true, true, true, pc, likB, likD, unused);
    const double pr = warp_sum(trend_prior_term(lane, mine));
    double adq[3];
    trend_adequacy(v, p, lane, adq);
    if (lane == 0) {
        if (lik) { lik[2 * s] = likB; lik[2 * s + 1] = likD; }
        if (prior) prior[s] = pr;
        if (out_hast) out_hast[s] = h;
        if (adequacy) { adequacy[3 * s] = adq[0]; adequacy[3 * s + 1] = adq[1]; adequacy[3 * s + 2] = adq[2]; }
    }
    if (out_params && lane < TR_NPAR) out_params[s * TR_NPAR + lane] = mine;
    if (rates)
        for (int j = lane; j < nb; j += 32) {
            const double lnT = v.tab[TR_LNT * nbp + j];
            rates[(size_t)s * 2 * nb + j] = trend_rate(p[0], p[2], p[4], lnT, const_b);
            rates[(size_t)s * 2 * nb + nb + j] = trend_rate(p[1], p[3], p[5], lnT, const_d);
        }
}

inline int trend_grid(int n, int& threads) {
    const int wpb = n <= 1024 ? 1 : 4;      // few chains: one warp per CTA so that they spread over all SMs
    threads = wpb * 32;
    return (n + wpb - 1) / wpb;
}

// per-parameter Bernoulli probabilities of the two proposal kinds (trend_rate.py:122-134)
void move_weights(int const_b, int const_d, double* f_mult, double* f_norm) {
    double m[TR_NPAR] = {1, 1, 0, 0, 1, 1}, nn[TR_NPAR] = {0, 0, 1, 1, 0, 0};
    if (const_b) { m[4] = 0; nn[2] = 0; }
    if (const_d) { m[5] = 0; nn[3] = 0; }
    double sm = 0, sn = 0;
    for (int k = 0; k < TR_NPAR; ++k) { sm += m[k]; sn += nn[k]; }
    for (int k = 0; k < TR_NPAR; ++k) { f_mult[k] = m[k] / sm; f_norm[k] = nn[k] / sn; }
}

}  // namespace

extern "C" int64_t lr_trend_record_doubles(int32_t n_bins) { return LR_TREND_REC_HEAD + 2 * (int64_t)n_bins; }

extern "C" int lr_trend_create(lr_handle_t h, int32_t n_rep, int32_t n_bins, const int64_t* d_sp, const int64_t* d_ex,
                               const double* d_br, const double* h_trend, int32_t const_birth, int32_t const_death,
                               int32_t n_chains, uint64_t seed, int64_t chain_id0, const int32_t* h_rep_of_chain,
                               void* stream, lr_trend_t* out) {
    LR_REQUIRE(h && d_sp && d_ex && d_br && h_trend && out, "lr_trend_create: null pointer");
    LR_REQUIRE(n_rep >= 1 && n_bins >= 1 && n_chains >= 1, "lr_trend_create: n_rep, n_bins, n_chains must be >= 1");
    LR_REQUIRE(chain_id0 >= 0 && chain_id0 + n_chains <= 0xffffffffll, "lr_trend_create: chain ids must fit 32 bits");
    if (const_birth && const_death) {
        lr_set_error("lr_trend_create: constant birth AND death rates leave the additive move without a parameter; the reference "
                     "stops in np.random.binomial(p = nan) on its first such move (trend_rate.py:129-134, :168)");
        return LR_ERR_UNSUPPORTED;
    }
    for (int j = 0; j < n_bins; ++j)
        LR_REQUIRE(h_trend[j] > 0.0 && h_trend[j] <= 1.0, "lr_trend_create: trend[%d] = %g outside (0, 1]; pass the min-max normalised trend "
                   "with zeros replaced by 1e-15 (trend_rate.py:65-68)", j, h_trend[j]);
    if (h_rep_of_chain)
        for (int i = 0; i < n_chains; ++i)
            LR_REQUIRE(h_rep_of_chain[i] >= 0 && h_rep_of_chain[i] < n_rep, "lr_trend_create: replicate of chain %d out of range", i);
    LR_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    lr_trend_t t = new lr_trend_s();
    memset(t, 0, sizeof(*t));
    t->last.s = st; t->last.valid = 1;
    t->h = h; t->n_rep = n_rep; t->n_bins = n_bins; t->nbp = (n_bins + 3) & ~3;
    t->const_b = const_birth != 0; t->const_d = const_death != 0; t->n_chains = n_chains; t->seed = seed;
    move_weights(t->const_b, t->const_d, t->f_mult, t->f_norm);
    const size_t tab_bytes = (size_t)n_rep * TR_ROWS * t->nbp * sizeof(double);
    cudaError_t e = cudaMallocAsync((void**)&t->tab, tab_bytes, st);
    if (e == cudaSuccess) e = cudaMallocAsync((void**)&t->cst, (size_t)n_rep * 2 * sizeof(double), st);
    if (e == cudaSuccess) e = cudaMallocAsync((void**)&t->st, (size_t)n_chains * sizeof(TrendChain), st);
    if (e != cudaSuccess) { lr_set_error("lr_trend_create: cudaMallocAsync: %s", cudaGetErrorString(e)); lr_trend_destroy(t); return LR_ERR_NOMEM; }
    // trend and replicate map through stream-ordered scratch
    double* d_trend = nullptr;
    int* d_rep = nullptr;
    e = cudaMallocAsync((void**)&d_trend, (size_t)n_bins * sizeof(double), st);
    if (e == cudaSuccess && h_rep_of_chain) e = cudaMallocAsync((void**)&d_rep, (size_t)n_chains * sizeof(int), st);
    if (e != cudaSuccess) { lr_set_error("lr_trend_create: cudaMallocAsync: %s", cudaGetErrorString(e)); lr_trend_destroy(t); return LR_ERR_NOMEM; }
    auto undo = [&]() { if (d_trend) cudaFreeAsync(d_trend, st); if (d_rep) cudaFreeAsync(d_rep, st); lr_trend_destroy(t); };
    LR_CUDA_CLEAN(cudaMemcpyAsync(d_trend, h_trend, (size_t)n_bins * sizeof(double), cudaMemcpyHostToDevice, st), undo());
    if (d_rep) LR_CUDA_CLEAN(cudaMemcpyAsync(d_rep, h_rep_of_chain, (size_t)n_chains * sizeof(int), cudaMemcpyHostToDevice, st), undo());
    k6_build_tables<<<n_rep, 128, 0, st>>>((const long long*)d_sp, (const long long*)d_ex, d_br, d_trend, n_bins, t->nbp, t->tab, t->cst);
    LR_CUDA_CLEAN(cudaGetLastError(), undo());
    int threads;
    const int blocks = trend_grid(n_chains, threads);
    k6_init_kernel<<<blocks, threads, 0, st>>>(t->st, n_chains, d_rep, chain_id0, t->tab, t->cst, n_bins, t->nbp, t->const_b, t->const_d);
    LR_CUDA_CLEAN(cudaGetLastError(), undo());
    h->launches += 2;
    LR_CUDA_CLEAN(cudaStreamSynchronize(st), undo());       // h_trend / h_rep_of_chain may be pageable: they are consumed when this returns
    cudaFreeAsync(d_trend, st);
    if (d_rep) cudaFreeAsync(d_rep, st);
    *out = t;
    return LR_OK;
}

extern "C" int lr_trend_create_host(lr_handle_t h, int32_t n_rep, int32_t n_bins, const int64_t* h_sp, const int64_t* h_ex,
                                    const double* h_br, const double* h_trend, int32_t const_birth, int32_t const_death,
                                    int32_t n_chains, uint64_t seed, int64_t chain_id0, const int32_t* h_rep_of_chain,
                                    lr_trend_t* out) {
    LR_REQUIRE(h && h_sp && h_ex && h_br, "lr_trend_create_host: null pointer");
    LR_REQUIRE(n_rep >= 1 && n_bins >= 1, "lr_trend_create_host: bad sizes");
    LR_CUDA(cudaSetDevice(h->device));
    const size_t cnt = (size_t)n_rep * n_bins;
    int rc = lr_ws_acquire(h, cnt * 24, h->stream);
    if (rc != LR_OK) return rc;
    int64_t* d_sp = (int64_t*)h->ws;
    int64_t* d_ex = d_sp + cnt;
    double* d_br = (double*)(d_ex + cnt);
    LR_CUDA(cudaMemcpyAsync(d_sp, h_sp, cnt * 8, cudaMemcpyHostToDevice, h->stream));
    LR_CUDA(cudaMemcpyAsync(d_ex, h_ex, cnt * 8, cudaMemcpyHostToDevice, h->stream));
    LR_CUDA(cudaMemcpyAsync(d_br, h_br, cnt * 8, cudaMemcpyHostToDevice, h->stream));
    return lr_trend_create(h, n_rep, n_bins, d_sp, d_ex, d_br, h_trend, const_birth, const_death, n_chains, seed, chain_id0,
                           h_rep_of_chain, h->stream, out);
}

extern "C" int lr_trend_destroy(lr_trend_t t) {
    if (!t) return LR_OK;
    cudaSetDevice(t->h->device);
    lr_order(t->h, &t->last, t->h->stream);
    if (t->tab) cudaFreeAsync(t->tab, t->h->stream);
    if (t->cst) cudaFreeAsync(t->cst, t->h->stream);
    if (t->st) cudaFreeAsync(t->st, t->h->stream);
    delete t;
    return LR_OK;
}

extern "C" int64_t lr_trend_records_per_run(lr_trend_t t, int64_t n_iter, int64_t sample_every) {
    if (!t || n_iter <= 0 || sample_every <= 0) return 0;
    long long it0 = 0;
    cudaSetDevice(t->h->device);
    lr_order(t->h, &t->last, t->h->stream);
    cudaStreamSynchronize(t->h->stream);
    if (cudaMemcpy(&it0, &t->st[0].it, sizeof(long long), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    const long long first = (it0 + sample_every - 1) / sample_every * sample_every;
    const long long it1 = it0 + n_iter;
    return first < it1 ? (it1 - 1 - first) / sample_every + 1 : 0;
}

extern "C" int lr_trend_run(lr_trend_t t, int64_t n_iter, int64_t sample_every, double* d_records, void* stream) {
    LR_REQUIRE(t != nullptr, "lr_trend_run: null chains");
    LR_REQUIRE(n_iter >= 0 && sample_every >= 0, "lr_trend_run: negative count");
    if (n_iter == 0) return LR_OK;
    lr_handle_t h = t->h;
    LR_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    { int rc0 = lr_order(h, &t->last, st); if (rc0 != LR_OK) return rc0; }
    TrendRun P;
    P.st = t->st; P.n_chains = t->n_chains; P.tab = t->tab; P.cst = t->cst; P.nb = t->n_bins; P.nbp = t->nbp;
    P.const_b = t->const_b; P.const_d = t->const_d; P.k0 = (uint32_t)t->seed; P.k1 = (uint32_t)(t->seed >> 32);
    P.n_iter = n_iter; P.sample_every = sample_every > 0 ? sample_every : (int64_t)1 << 62;
    P.records = sample_every > 0 ? d_records : nullptr;
    P.rec_doubles = (int)lr_trend_record_doubles(t->n_bins);
    for (int k = 0; k < TR_NPAR; ++k) { P.fd_mult[k] = t->f_mult[k]; P.fd_norm[k] = t->f_norm[k]; }
    int threads;
    const int blocks = trend_grid(t->n_chains, threads);
    // wide build (one warp per canonical group of bins) when one warp per chain would leave most of the GPU idle: measured on
    // B200 at 200 bins, M it/s one-warp / wide: 256 chains 82 / 141, 2048 chains 454 / 213.  Same chains either way.
    int W = trend_groups(t->n_bins);
    if ((long long)t->n_chains * W > (long long)h->sm_count * 8) W = 1;
    { const char* e = getenv("LR_TREND_WIDE"); if (e) W = atoi(e) ? trend_groups(t->n_bins) : 1; }         // development override
    const size_t smem = (size_t)4 * t->nbp * sizeof(double);
    if (W > 1 && smem > 40 * 1024) W = 1;
    if (W == 4) k6_trend_wide_kernel<4><<<t->n_chains, 128, smem, st>>>(P);
    else if (W == 2) k6_trend_wide_kernel<2><<<t->n_chains, 64, smem, st>>>(P);
    else k6_trend_kernel<<<blocks, threads, 0, st>>>(P);
    LR_CUDA(cudaGetLastError());
    h->launches += 1;
    return LR_OK;
}

extern "C" int lr_trend_run_host(lr_trend_t t, int64_t n_iter, int64_t sample_every, double* h_records) {
    LR_REQUIRE(t != nullptr, "lr_trend_run_host: null chains");
    lr_handle_t h = t->h;
    const int64_t nrec = (h_records && sample_every > 0) ? lr_trend_records_per_run(t, n_iter, sample_every) : 0;
    LR_REQUIRE(nrec >= 0, "lr_trend_run_host: could not read the iteration counter");
    const size_t bytes = (size_t)nrec * t->n_chains * lr_trend_record_doubles(t->n_bins) * sizeof(double);
    double* d_rec = nullptr;
    if (bytes) {
        int rc = lr_ws_acquire(h, bytes, h->stream);
        if (rc != LR_OK) return rc;
        d_rec = (double*)h->ws;
    }
    int rc = lr_trend_run(t, n_iter, bytes ? sample_every : 0, d_rec, h->stream);
    if (rc != LR_OK) return rc;
    if (bytes) LR_CUDA(cudaMemcpyAsync(h_records, d_rec, bytes, cudaMemcpyDeviceToHost, h->stream));
    LR_CUDA(cudaStreamSynchronize(h->stream));
    return LR_OK;
}

extern "C" int lr_trend_eval_host(lr_trend_t t, int32_t n, const int32_t* rep, const double* params, const int32_t* kind,
                                  const int32_t* on, const double* draw, double* out_params, double* out_hast, double* lik,
                                  double* prior, double* rates, double* adequacy) {
    LR_REQUIRE(t && params, "lr_trend_eval_host: null pointer");
    LR_REQUIRE(n >= 1, "lr_trend_eval_host: n must be >= 1");
    LR_REQUIRE(!kind || (on && draw), "lr_trend_eval_host: kind needs on and draw");
    if (rep)
        for (int i = 0; i < n; ++i) LR_REQUIRE(rep[i] >= 0 && rep[i] < t->n_rep, "lr_trend_eval_host: replicate of state %d out of range", i);
    lr_handle_t h = t->h;
    LR_CUDA(cudaSetDevice(h->device));
    { int rc0 = lr_order(t->h, &t->last, t->h->stream); if (rc0 != LR_OK) return rc0; }
    const int nb = t->n_bins;
    // workspace layout (doubles unless noted)
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    const size_t o_par = take((size_t)n * TR_NPAR * 8), o_rep = take((size_t)n * 4), o_kind = take((size_t)n * 4),
                 o_on = take((size_t)n * TR_NPAR * 4), o_draw = take((size_t)n * TR_NPAR * 8), o_np = take((size_t)n * TR_NPAR * 8),
                 o_h = take((size_t)n * 8), o_lik = take((size_t)n * 16), o_pr = take((size_t)n * 8),
                 o_rates = take((size_t)n * 2 * nb * 8), o_adq = take((size_t)n * 24);
    int rc = lr_ws_acquire(h, off, h->stream);
    if (rc != LR_OK) return rc;
    char* W = (char*)h->ws;
    cudaStream_t st = h->stream;
    LR_CUDA(cudaMemcpyAsync(W + o_par, params, (size_t)n * TR_NPAR * 8, cudaMemcpyHostToDevice, st));
    if (rep) LR_CUDA(cudaMemcpyAsync(W + o_rep, rep, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    if (kind) {
        LR_CUDA(cudaMemcpyAsync(W + o_kind, kind, (size_t)n * 4, cudaMemcpyHostToDevice, st));
        LR_CUDA(cudaMemcpyAsync(W + o_on, on, (size_t)n * TR_NPAR * 4, cudaMemcpyHostToDevice, st));
        LR_CUDA(cudaMemcpyAsync(W + o_draw, draw, (size_t)n * TR_NPAR * 8, cudaMemcpyHostToDevice, st));
    }
    k6_eval_kernel<<<(n + 3) / 4, 128, 0, st>>>(t->tab, t->cst, nb, t->nbp, t->const_b, t->const_d, n, rep ? (const int*)(W + o_rep) : nullptr,
                                                (const double*)(W + o_par), kind ? (const int*)(W + o_kind) : nullptr,
                                                (const int*)(W + o_on), (const double*)(W + o_draw), (double*)(W + o_np), (double*)(W + o_h),
                                                (double*)(W + o_lik), (double*)(W + o_pr), (double*)(W + o_rates), (double*)(W + o_adq));
    LR_CUDA(cudaGetLastError());
    h->launches += 1;
    if (out_params) LR_CUDA(cudaMemcpyAsync(out_params, W + o_np, (size_t)n * TR_NPAR * 8, cudaMemcpyDeviceToHost, st));
    if (out_hast) LR_CUDA(cudaMemcpyAsync(out_hast, W + o_h, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    if (lik) LR_CUDA(cudaMemcpyAsync(lik, W + o_lik, (size_t)n * 16, cudaMemcpyDeviceToHost, st));
    if (prior) LR_CUDA(cudaMemcpyAsync(prior, W + o_pr, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    if (rates) LR_CUDA(cudaMemcpyAsync(rates, W + o_rates, (size_t)n * 2 * nb * 8, cudaMemcpyDeviceToHost, st));
    if (adequacy) LR_CUDA(cudaMemcpyAsync(adequacy, W + o_adq, (size_t)n * 24, cudaMemcpyDeviceToHost, st));
    LR_CUDA(cudaStreamSynchronize(st));
    return LR_OK;
}

extern "C" int lr_trend_state_host(lr_trend_t t, double* h_state) {
    LR_REQUIRE(t && h_state, "lr_trend_state_host: null pointer");
    LR_CUDA(cudaSetDevice(t->h->device));
    { int rc0 = lr_order(t->h, &t->last, t->h->stream); if (rc0 != LR_OK) return rc0; }
    LR_CUDA(cudaStreamSynchronize(t->h->stream));
    // [n_chains][LR_TREND_STATE_DOUBLES]: parameters, likB, likD, prior, then iteration and accepted as doubles
    TrendChain* tmp = new TrendChain[t->n_chains];
    cudaError_t e = cudaMemcpy(tmp, t->st, (size_t)t->n_chains * sizeof(TrendChain), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { delete[] tmp; lr_set_error("lr_trend_state_host: %s", cudaGetErrorString(e)); return LR_ERR_CUDA; }
    for (int c = 0; c < t->n_chains; ++c) {
        double* o = h_state + (size_t)c * LR_TREND_STATE_DOUBLES;
        for (int k = 0; k < TR_NPAR; ++k) o[k] = tmp[c].p[k];
        o[6] = tmp[c].likB; o[7] = tmp[c].likD; o[8] = tmp[c].prior; o[9] = (double)tmp[c].it; o[10] = (double)tmp[c].accepted;
        o[11] = (double)tmp[c].rep;
    }
    delete[] tmp;
    return LR_OK;
}
