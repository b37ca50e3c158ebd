"""CPU oracle for the LiteRate RJMCMC birth-death hot path.  TEST INFRASTRUCTURE ONLY.

This module is a plain NumPy restatement of the algorithm in the reference script
``LiteRateForward.py`` (which cannot be imported: it parses ``sys.argv`` and runs a chain
at module level).  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it, and only as the checker or
the CPU baseline -- never on the product path.

Parity pin: ``oracle/make_golden.py`` runs the *unmodified* reference in a subprocess in
the build container and the oracle chain (same seed, same legacy ``np.random`` draw
order) reproduces its four log files byte for byte; the resulting fixtures are committed
under ``tests/golden/`` and re-checked by the CPU test-suite.  The reference itself ships
no tests or golden vectors for this path (the known-answer example of tutorial 2 is the
only one; it is checked too).

All ``file:line`` citations are relative to the reference checkout.
"""
from __future__ import annotations

import csv
import io
import math
import os
from dataclasses import dataclass, field

import numpy as np

# --------------------------------------------------------------------------------------
# constants set at module level by the reference (LiteRateForward.py:585-590)
# --------------------------------------------------------------------------------------
SHAPE_BETA_RJ = 10.0      # :586
MIN_ALLOWED_T = 1         # :587
GAMMA_SHAPE = 2.0         # :588
HP_GAMMA_SHAPE = 1.2      # :589
HP_GAMMA_RATE = 0.1       # :590
MODEL_SUFFIX = {0: "_BD", 1: "_ID", 2: "_BDk", 3: "_BDd"}   # :422-427

MCMC_HEADER = ["it", "posterior", "likelihood", "prior", "lambda_avg", "mu_avg", "K_l", "K_m",
               "root_age", "death_age", "gamma_rate_hp_BI", "gamma_rate_hp_D",
               "poisson_rate_hp"]                                   # :496-502
ADEQUACY_HEADER = ["corr_coeff", "rsquared", "gelman_r2"]           # :498


# --------------------------------------------------------------------------------------
# L1: input parsing (LiteRateForward.py:439-476)
# --------------------------------------------------------------------------------------
@dataclass
class Lineages:
    ts: np.ndarray
    te: np.ndarray
    start_time: float
    end_time: float
    true_root_age: float


def read_lineages(path, TBP=False, rev_se=0, first_year=-1, last_year=-1, death_jitter=0.5):
    """TSV -> (ts, te) following :440-474.

    ``-first_year`` raises IndexError in the reference whenever it filters anything
    (:460-461 masks ``te_years`` with the already filtered ``ts_years``); here the
    intended order of ``literate_library.parse_ts_te`` (literate_library.py:216-222) is
    used, as SURVEY Appendix A-12 prescribes.
    """
    tbl = np.genfromtxt(path, skip_header=1)
    if tbl.ndim == 1:
        tbl = tbl[None, :]
    if tbl.shape[1] == 4:           # deprecated (clade, id, ts, te) layout, :441-444
        ts_y, te_y = tbl[:, 2], tbl[:, 3]
    elif rev_se:                    # :446-448
        ts_y, te_y = tbl[:, 2], tbl[:, 1]
    else:                           # :450-451
        ts_y, te_y = tbl[:, 1], tbl[:, 2]
    if TBP:                         # :453-456
        root = np.max(ts_y)
        ts, te = root - ts_y, root - te_y
    else:                           # :457-468
        root = 0
        if first_year != -1:
            keep = ts_y >= first_year
            ts_y, te_y = ts_y[keep], te_y[keep]
        if last_year != -1:
            keep = ts_y <= last_year
            ts, te = ts_y[keep], te_y[keep] + 0.0
            te[te > last_year] = last_year
        else:
            ts, te = ts_y, te_y
    te = te + death_jitter          # :471
    return Lineages(np.ascontiguousarray(ts, dtype=np.float64),
                    np.ascontiguousarray(te, dtype=np.float64),
                    np.min(ts), np.max(te), root)


# --------------------------------------------------------------------------------------
# L2: lineages -> per-bin sufficient statistics (LiteRateForward.py:111-123, :514-549)
# --------------------------------------------------------------------------------------
def time_at_risk(ts, te, t0, t1):
    """get_br (:111-116): sum of the positive parts of min(te,t1)-max(ts,t0)."""
    lo = np.where(ts < t0, t0, ts)
    hi = np.where(te > t1, t1, te)
    d = hi - lo
    return np.sum(d[d > 0])


def events_in_bin(ts, te, t0, t1):
    """precompute_events (:118-123): births ts in [t0,t1), deaths te in (t0,t1]."""
    n_sp = int(np.count_nonzero((ts >= t0) & (ts < t1)))
    n_ex = int(np.count_nonzero((te > t0) & (te <= t1)))
    return n_sp, n_ex, time_at_risk(ts, te, t0, t1)


def events_in_bin_asref(ts, te, t0, t1):
    """Same three numbers as :func:`events_in_bin`, computed with the SAME NumPy operations the
    reference issues (:111-123): index sets through ``nonzero`` + ``np.intersect1d`` (two sorts) for
    the counts, two array copies + two masked stores for the clip.  ~20x slower than
    :func:`events_in_bin`; used where the reference's *cost* is what is measured (bench.py's CPU
    baseline), and checked against :func:`events_in_bin` in the CPU tests."""
    born = np.intersect1d((ts >= t0).nonzero()[0], (ts < t1).nonzero()[0])
    died = np.intersect1d((te > t0).nonzero()[0], (te <= t1).nonzero()[0])
    lo, hi = ts + 0., te + 0.
    lo[lo < t0] = t0
    hi[hi > t1] = t1
    span = hi - lo
    return len(born), len(died), np.sum(span[span > 0])


def bin_range(ts, te):
    """range(int(min ts), int(max te)) of :519."""
    return range(int(np.min(ts)), int(np.max(te)))


@dataclass
class BinStats:
    first_bin: int
    sp: np.ndarray            # int64 [n_bins]
    ex: np.ndarray            # int64 [n_bins]
    br: np.ndarray            # float64 [n_bins]
    ex_dead: np.ndarray = None  # model_BDI 3 only (:529-549)
    br_dead: np.ndarray = None

    @property
    def n_bins(self):
        return len(self.sp)


def bin_stats(ts, te, only_dead=False, end_time=None):
    """One (sp, ex, br) triple per unit bin, :519-523; extinct-only variant :529-549."""
    rng = bin_range(ts, te)
    sp, ex, br = [], [], []
    for i in rng:
        a, b, c = events_in_bin(ts, te, i, i + 1)
        sp.append(a), ex.append(b), br.append(c)
    out = BinStats(rng.start, np.array(sp, dtype=np.int64), np.array(ex, dtype=np.int64),
                   np.array(br, dtype=np.float64))
    if only_dead:
        if end_time is None:
            end_time = np.max(te)
        keep = te < end_time                      # :531-532
        s, e = ts[keep], te[keep]
        exd, brd = [], []
        for i in rng:
            exd.append(int(np.count_nonzero((e > i) & (e <= i + 1))))
            brd.append(time_at_risk(s, e, i, i + 1))
        out.ex_dead = np.array(exd, dtype=np.int64)
        out.br_dead = np.array(brd, dtype=np.float64)
    return out


def bin_stats_fast(ts, te, only_dead=False, end_time=None):
    """Same numbers as :func:`bin_stats` for data whose fractional parts are multiples
    of 2^-20 (every shipped data set: integer years + .5 jitter), in O(N) instead of
    O(N * n_bins).  Used only to build expectations at the full 1M/100M-lineage sizes the
    per-bin loop cannot finish in seconds; it is itself checked against
    :func:`bin_stats` in the CPU tests.  Identity: SURVEY 7.3.
    """
    s0 = int(np.min(ts))
    nb = int(np.max(te)) - s0

    def one(ts, te):
        a = np.floor(ts).astype(np.int64) - s0
        b = np.ceil(te).astype(np.int64) - 1 - s0
        live = te > ts
        okb = (a >= 0) & (a < nb)
        okd = (b >= 0) & (b < nb)
        sp = np.bincount(a[okb], minlength=nb)[:nb].astype(np.int64)
        ex = np.bincount(b[okd], minlength=nb)[:nb].astype(np.int64)
        br = np.zeros(nb)
        # dense difference array on the clipped interval, exact for dyadic fractions
        lo = np.clip(ts[live], s0, s0 + nb)
        hi = np.clip(te[live], s0, s0 + nb)
        keep = hi > lo
        lo, hi = lo[keep], hi[keep]
        ia = np.minimum(np.floor(lo).astype(np.int64) - s0, nb - 1)
        ib = np.minimum(np.ceil(hi).astype(np.int64) - 1 - s0, nb - 1)
        same = ia == ib
        np.add.at(br, ia[same], (hi - lo)[same])
        d = ~same
        np.add.at(br, ia[d], (s0 + ia[d] + 1) - lo[d])
        np.add.at(br, ib[d], hi[d] - (s0 + ib[d]))
        full = np.zeros(nb + 1)
        np.add.at(full, ia[d] + 1, 1.0)
        np.add.at(full, ib[d], -1.0)
        br += np.cumsum(full)[:nb]
        return sp, ex, br

    sp, ex, br = one(ts, te)
    out = BinStats(s0, sp, ex, br)
    if only_dead:
        if end_time is None:
            end_time = np.max(te)
        keep = te < end_time
        _, exd, brd = one(ts[keep], te[keep])
        out.ex_dead, out.br_dead = exd, brd
    return out


def write_div_log(fh, stats):
    """div.log (:558-564): '\\n' after the header, csv.writer's '\\r\\n' after every row."""
    fh.write("sp_events\tex_events\tbr_length\n")
    w = csv.writer(fh, delimiter="\t")
    for row in zip(stats.sp.tolist(), stats.ex.tolist(), [np.float64(x) for x in stats.br]):
        w.writerow(row)


# --------------------------------------------------------------------------------------
# L3: likelihood and priors on a (rates, shift-times) state
# --------------------------------------------------------------------------------------
def rate_index(times, n_bins):
    """get_rate_index (:125-135).  Callers pass np.floor(times) except for the initial
    two-element state (:224-225, :262, :272, :277-278)."""
    if len(times) == 2:
        return np.zeros(n_bins, dtype=int)
    t = np.round(times + 0)
    reps = np.abs(np.diff(t)).astype(int)
    return np.repeat(np.arange(len(t) - 1), reps)


def loglik_bdi(lam_bins, mu_bins, stats, model_BDI):
    """BDI_partial_lik (:150-162) with Tk == 1 (:574)."""
    k = stats.br
    m = k > 0
    L = lam_bins * (1 - model_BDI)
    I = lam_bins * model_BDI
    kk = k[m]
    terms = stats.sp[m] * np.log(kk * L[m] + I[m]) + stats.ex[m] * np.log(mu_bins[m] * kk) \
        - 1.0 * (kk * (L[m] + mu_bins[m]) + I[m])
    return np.sum(terms)


def loglik_keiding(lam_bins, mu_bins, stats, only_dead):
    """BD_lik_Keiding (:137-148)."""
    b = np.sum(np.log(lam_bins) * stats.sp - lam_bins * stats.br)
    if only_dead:
        d = np.sum(np.log(mu_bins) * stats.ex_dead - mu_bins * stats.br_dead)
    else:
        d = np.sum(np.log(mu_bins) * stats.ex - mu_bins * stats.br)
    return b + d


def loglik_state(L, M, timesL, timesM, stats, model_BDI, floor_times=True):
    """Likelihood of a full state the way the loop evaluates it (:262, :272, :277-278, :306)."""
    nb = stats.n_bins
    iL = rate_index(np.floor(timesL) if floor_times and len(timesL) > 2 else timesL, nb)
    iM = rate_index(np.floor(timesM) if floor_times and len(timesM) > 2 else timesM, nb)
    lam, mu = np.asarray(L)[iL], np.asarray(M)[iM]
    if model_BDI <= 1:
        return loglik_bdi(lam, mu, stats, model_BDI)
    return loglik_keiding(lam, mu, stats, model_BDI == 3)


def ln_gamma_pdf(x, shape, rate):
    """scipy.stats.gamma.logpdf(x, shape, scale=1/rate) in closed form (:201-202)."""
    x = np.asarray(x, dtype=np.float64)
    return shape * math.log(rate) - math.lgamma(shape) + (shape - 1.0) * np.log(x) - rate * x


def ln_sym_beta_pdf(u, a):
    """scipy.stats.beta.logpdf(u, a, a) in closed form (:22-23)."""
    return (a - 1.0) * (math.log(u) + math.log1p(-u)) - (2.0 * math.lgamma(a) - math.lgamma(2.0 * a))


def poisson_prior(k, rate):
    """Poisson_prior (:198-199)."""
    return k * np.log(rate) - rate - np.sum(np.log(np.arange(1, k + 1)))


def rates_prior(x, shape=2, rate=2, exact_scipy=False):
    """prior_gamma (:201-202)."""
    if exact_scipy:
        import scipy.stats
        return np.sum(scipy.stats.gamma.logpdf(x, shape, scale=1. / rate, loc=0))
    return np.sum(ln_gamma_pdf(x, shape, rate))


def state_prior(L, M, gamma_rate, span, prior_poi, exact_scipy=False):
    """prior as assembled at :296-303 (prior_poi is priorPoi or the stale priorPoiA)."""
    p = rates_prior(L, GAMMA_SHAPE, gamma_rate[0], exact_scipy) + \
        rates_prior(M, GAMMA_SHAPE, gamma_rate[1], exact_scipy)
    p += -np.log(span) * (len(L) - 1 + len(M) - 1)
    return p + prior_poi


def adequacy(emp_birth, emp_death, est_birth, est_death):
    """calculate_r_squared (literate_library.py:268-279)."""
    x = np.concatenate([emp_birth, emp_death])
    y = np.concatenate([est_birth, est_death])
    sol = np.linalg.lstsq(np.vstack(x), y, rcond=None)
    coeff = sol[0][0]
    r2 = 1 - sol[1][0] / np.sum(y ** 2)
    fit = coeff * x
    res = y - fit
    vf = np.var(fit, ddof=1)
    return coeff, r2, vf / (vf + np.var(res, ddof=1))


def adequacy_closed_form(emp_birth, emp_death, est_birth, est_death):
    """Same three numbers without lstsq: one-regressor least squares through the origin."""
    x = np.concatenate([emp_birth, emp_death])
    y = np.concatenate([est_birth, est_death])
    sxx, sxy, syy = np.sum(x * x), np.sum(x * y), np.sum(y * y)
    coeff = sxy / sxx
    ssres = np.sum((y - coeff * x) ** 2)
    fit = coeff * x
    res = y - fit
    vf = np.var(fit, ddof=1)
    return coeff, 1 - ssres / syy, vf / (vf + np.var(res, ddof=1))


# --------------------------------------------------------------------------------------
# L4: proposals.  Each consumes the legacy global np.random stream in the reference's order
# --------------------------------------------------------------------------------------
def _sym_beta_logpdf(u, exact_scipy):
    if exact_scipy:
        import scipy.stats
        return scipy.stats.beta.logpdf(u, SHAPE_BETA_RJ, SHAPE_BETA_RJ)
    return ln_sym_beta_pdf(u, SHAPE_BETA_RJ)


def rate_multiplier_given(q, touched, u, d=1.1):
    """update_multiplier_freq (:165-176) for a given Bernoulli mask and uniforms."""
    m = np.exp(2 * np.log(d) * (np.asarray(u, float) - .5))
    m[np.asarray(touched) == 0] = 1.
    return q * m, np.sum(np.log(m))


def propose_rate_multiplier(q, f, d=1.1):
    """update_multiplier_freq (:165-176) with the reference's draws."""
    shape = np.shape(q)
    touched = np.random.binomial(1, f, shape)
    u = np.random.uniform(0, 1, shape)
    return rate_multiplier_given(q, touched, u, d)


def propose_move_shift(times, start_time, end_time):
    """update_times (:188-195) + update_sliding_win (:178-186): two draws are consumed and the
    proposal is, by construction of :184-185, the current vector again."""
    out = times + 0.
    j = np.random.choice(range(1, len(times) - 1))
    np.random.random()
    return np.sort(out)


def add_shift_given(rates, times, i, t_new, u, exact_scipy=False):
    """add_shift_RJ_weighted_mean (:29-47) for given draws: segment i, new shift time t_new, Beta(10,10) variate u."""
    gap = times[i + 1] - times[i]
    times_new = np.sort(np.array(list(times) + [t_new]))
    ta, tb = times[i], times[i + 1]
    p1 = (ta - t_new) / (ta - tb)
    p2 = (t_new - tb) / (ta - tb)
    r = rates[i]
    r1 = np.exp(np.log(r) - p2 * np.log((1 - u) / u))
    r2 = np.exp(np.log(r) + p1 * np.log((1 - u) / u))
    rates_new = np.insert(rates, i + 1, r2)
    rates_new[i] = r1
    log_q = np.log(abs(gap)) - _sym_beta_logpdf(u, exact_scipy)
    jac = 2 * np.log(r1 + r2) - np.log(r)
    return rates_new, times_new, log_q + jac


def propose_add_shift(rates, times, exact_scipy=False):
    """add_shift_RJ_weighted_mean (:29-47) with the reference's draws from the global stream, in its order."""
    gaps = np.diff(times)
    i = np.random.choice(range(len(gaps)))
    t_new = times[i] + np.random.uniform(0, gaps[i])
    u = np.random.beta(SHAPE_BETA_RJ, SHAPE_BETA_RJ)
    return add_shift_given(rates, times, i, t_new, u, exact_scipy)


def remove_shift_given(rates, times, j, exact_scipy=False):
    """remove_shift_RJ_weighted_mean (:49-69) for a given interior shift j (1 <= j <= len(times) - 2)."""
    t_rm, ta, tb = times[j], times[j - 1], times[j + 1]
    span = abs(tb - ta)
    times_new = times[times != t_rm]
    p1 = (ta - t_rm) / (ta - tb)
    p2 = (t_rm - tb) / (ta - tb)
    ra, rb = rates[j - 1], rates[j]
    merged = np.exp(p1 * np.log(ra) + p2 * np.log(rb))
    rates_new = rates[rates != rates[j]]
    rates_new[j - 1] = merged
    u = 1. / (1 + rb / ra)
    log_q = -np.log(span) + _sym_beta_logpdf(u, exact_scipy)
    jac = np.log(merged) - (2 * np.log(ra + rb))
    return rates_new, times_new, log_q + jac


def propose_remove_shift(rates, times, exact_scipy=False):
    """remove_shift_RJ_weighted_mean (:49-69) with the reference's draw."""
    j = np.random.choice(range(1, len(times) - 1))
    return remove_shift_given(rates, times, j, exact_scipy)


def propose_rj(L, M, tL, tM, sample_shift_mu, exact_scipy=False):
    """RJMCMC (:71-97)."""
    r = np.random.random(2)
    nL, ntL, qL = L, tL, 0
    nM, ntM, qM = M, tM, 0
    if r[0] > sample_shift_mu:
        if r[1] > 0.5:
            nL, ntL, qL = propose_add_shift(L, tL, exact_scipy)
        elif len(L) > 1:
            nL, ntL, qL = propose_remove_shift(L, tL, exact_scipy)
        birth_side = 1
    else:
        if r[1] > 0.5:
            nM, ntM, qM = propose_add_shift(M, tM, exact_scipy)
        elif len(M) > 1:
            nM, ntM, qM = propose_remove_shift(M, tM, exact_scipy)
        birth_side = 0
    return nL, ntL, nM, ntM, qL + qM, birth_side


def gibbs_poisson_rate(k_l, k_m):
    """get_post_rj_HP (:99-108)."""
    return np.random.gamma(2. + k_l + k_m, 1. / (1. + 2))


def gibbs_gamma_rate(rates):
    """get_rate_HP (:210-213)."""
    rates = np.array(list(rates))
    return np.random.gamma(shape=HP_GAMMA_SHAPE + GAMMA_SHAPE * len(rates),
                           scale=1. / (HP_GAMMA_RATE + np.sum(rates)))


# --------------------------------------------------------------------------------------
# L4/L5: the chain (runMCMC, :216-373) and its log rows
# --------------------------------------------------------------------------------------
@dataclass
class ChainConfig:
    n_iterations: int = 10000000
    s_freq: int = 1000
    model_BDI: int = 0
    const_rates: int = 0
    const_death_rate: int = 0
    use_rate_HP: int = 1
    Poisson_prior: float = 0
    calc_adequacy: int = 1
    update_fraction: float = 0.75
    pyrate_output: bool = False
    exact_scipy: bool = False


@dataclass
class ChainLogs:
    mcmc: io.StringIO = field(default_factory=io.StringIO)
    sp: io.StringIO = field(default_factory=io.StringIO)
    ex: io.StringIO = field(default_factory=io.StringIO)
    n_lik_evals: int = 0
    n_accept: int = 0
    final_state: tuple = None


def mcmc_header(calc_adequacy):
    cols = MCMC_HEADER + (ADEQUACY_HEADER if calc_adequacy else [])
    return "\t".join(cols) + "\n"


def run_chain(lin: Lineages, stats: BinStats, cfg: ChainConfig, seed, logs: ChainLogs = None, lik_fn=None, emp=None):
    """runMCMC (:216-373) with the initial draw of :580-583 and the seeding of :405-409.

    Uses the legacy global ``np.random`` stream in exactly the reference's draw order, so
    that with the same ``-seed`` the log rows are the reference's log rows.
    """
    if logs is None:
        logs = ChainLogs()
    np.random.seed(seed)
    np.random.seed(seed)
    nb = stats.n_bins
    only_dead = cfg.model_BDI == 3
    start_time, end_time = lin.start_time, lin.end_time
    span = end_time - start_time

    def lik_of(L, iL, M, iM):
        if lik_fn is not None:             # another likelihood on the same per-bin rate vectors (oracle/proportion_oracle.py)
            return lik_fn(L[iL], M[iM])
        if cfg.model_BDI <= 1:
            return loglik_bdi(L[iL], M[iM], stats, cfg.model_BDI)
        return loglik_keiding(L[iL], M[iM], stats, only_dead)

    if cfg.calc_adequacy:
        with np.errstate(divide="ignore", invalid="ignore"):
            emp_b, emp_d = emp if emp is not None else (stats.sp / stats.br, stats.ex / stats.br)     # literate_library.py:260-266

    L_acc = np.random.gamma(2, 2, 1)        # :580
    M_acc = np.random.gamma(2, 2, 1)        # :581
    tLA = np.array([start_time, end_time])  # :582
    tMA = np.array([start_time, end_time])  # :583

    poi = 1 if cfg.Poisson_prior == 0 else cfg.Poisson_prior     # :220-221
    g_rate = [1., 1.]                                             # :222
    iLA, iMA = rate_index(tLA, nb), rate_index(tMA, nb)
    likA = lik_of(L_acc, iLA, M_acc, iMA)
    priorA = rates_prior(L_acc, exact_scipy=cfg.exact_scipy) + rates_prior(M_acc, exact_scipy=cfg.exact_scipy)  # :227
    priorA += -np.log(span) * (len(L_acc) - 1 + len(M_acc) - 1)
    poiA = poisson_prior(len(L_acc), poi) + poisson_prior(len(M_acc), poi)
    priorA += poiA

    if cfg.const_death_rate:                # :243-252
        shift_mu, b_freq, d_freq = 0, 0.7, 0.8
        fL, fM = cfg.update_fraction, 1
    else:
        shift_mu, b_freq, d_freq = 0.5, 0.4, 0.8
        fL, fM = cfg.update_fraction, cfg.update_fraction

    it = 0
    while it < cfg.n_iterations:
        r = np.random.random(2)
        L, tL = L_acc + 0, tLA + 0
        M, tM = M_acc + 0, tMA + 0
        iL, iM = iLA, iMA
        hasting, gibbs, poi_new = 0, 0, 0
        if r[0] < b_freq:
            if r[1] < .5 or len(L_acc) == 1:
                L, hasting = propose_rate_multiplier(L_acc, fL)
            else:
                tL = propose_move_shift(tLA, start_time, end_time)
                iL = rate_index(np.floor(tL), nb)
        elif r[0] < d_freq:
            if r[1] < .5 or len(M_acc) == 1:
                M, hasting = propose_rate_multiplier(M_acc, fM)
            else:
                tM = propose_move_shift(tMA, start_time, end_time)
                iM = rate_index(np.floor(tM), nb)
        elif r[0] < 0.999 and cfg.const_rates == 0:
            L, tL, M, tM, hasting, birth_side = propose_rj(L_acc, M_acc, tLA, tMA, shift_mu, cfg.exact_scipy)
            if birth_side == 1:
                iL = rate_index(np.floor(tL), nb)
            else:
                iM = rate_index(np.floor(tM), nb)
            poi_new = poisson_prior(len(L), poi) + poisson_prior(len(M), poi)
        else:
            if cfg.Poisson_prior == 0:
                poi = gibbs_poisson_rate(len(L_acc), len(M_acc))
            if cfg.use_rate_HP:
                g_rate = [gibbs_gamma_rate(L_acc), gibbs_gamma_rate(M_acc)]
            gibbs = 1

        if min(abs(np.diff(tL))) <= MIN_ALLOWED_T or min(abs(np.diff(tM))) <= MIN_ALLOWED_T:   # :290
            prior, lik = -np.inf, -np.inf
        else:
            if poi_new == 0:
                poi_new = poiA
            prior = state_prior(L, M, g_rate, span, poi_new, cfg.exact_scipy)
            if gibbs == 0:
                lik = lik_of(L, iL, M, iM)
                logs.n_lik_evals += 1
            else:
                lik = likA

        with np.errstate(divide="ignore"):
            log_u = np.log(np.random.random())
        if lik - likA + prior - priorA + hasting >= log_u or gibbs == 1:     # :313
            L_acc, M_acc, tLA, tMA = L, M, tL, tM
            likA, priorA = lik, prior
            iLA, iMA = iL, iM
            poiA = poi_new
            logs.n_accept += 1

        if it % cfg.s_freq == 0:           # :321-359
            if cfg.pyrate_output:
                root = lin.true_root_age
                head = [it, likA + priorA, likA, priorA, np.mean(L_acc), np.mean(M_acc),
                        len(L_acc), len(M_acc), root, root - np.max(tLA)]
                sp_t, ex_t = root - tLA[1:-1], root - tMA[1:-1]
            else:
                head = [it, likA + priorA, likA, priorA, np.mean(L_acc), np.mean(M_acc),
                        len(L_acc), len(M_acc), start_time, end_time]
                sp_t, ex_t = tLA[1:-1], tMA[1:-1]
            row = head + g_rate + [poi]
            if cfg.calc_adequacy:
                with np.errstate(divide="ignore", invalid="ignore"):
                    row += list(adequacy(emp_b, emp_d, L_acc[iLA], M_acc[iMA]))
            logs.mcmc.write("\t".join(map(str, row)) + "\n")
            logs.sp.write("\t".join(map(str, list(L_acc) + list(sp_t))) + "\n")
            logs.ex.write("\t".join(map(str, list(M_acc) + list(ex_t))) + "\n")
        it += 1
    logs.final_state = (L_acc, M_acc, tLA, tMA, likA, priorA, poiA, g_rate, poi)
    return logs


def run_reference_style(path, out_dir, seed, cfg: ChainConfig, TBP=False, death_jitter=0.5, out=""):
    """End to end: parse -> bin -> div.log -> chain -> the four log files, named as :486-512."""
    lin = read_lineages(path, TBP=TBP, death_jitter=death_jitter)
    stats = bin_stats(lin.ts, lin.te, only_dead=cfg.model_BDI == 3, end_time=lin.end_time)
    stem = os.path.splitext(os.path.basename(path))[0] + MODEL_SUFFIX[cfg.model_BDI] + out
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, stem + "_div.log"), "w", newline="") as fh:
        write_div_log(fh, stats)
    logs = ChainLogs()
    logs.mcmc.write(mcmc_header(cfg.calc_adequacy))
    run_chain(lin, stats, cfg, seed, logs)
    for tag, buf in (("mcmc", logs.mcmc), ("sp_rates", logs.sp), ("ex_rates", logs.ex)):
        with open(os.path.join(out_dir, f"{stem}_{tag}.log"), "w") as fh:
            fh.write(buf.getvalue())
    return lin, stats, logs


# --------------------------------------------------------------------------------------
# posterior summaries (plotRJforward.v3.py:92-139, :292-305) -- used by the chain parity tests
# --------------------------------------------------------------------------------------
def k_pmf(k_column, burnin=0.2):
    """get_K_values (plotRJforward.v3.py:292-305): unique/counts after burn-in."""
    n = len(k_column)
    b = min(int(burnin * n), int(0.9 * n))
    vals, counts = np.unique(np.asarray(k_column)[b:], return_counts=True)
    return dict(zip(vals.astype(int).tolist(), counts.tolist()))


def marginal_rates(rows, start_age, end_age, burnin=0.2):
    """get_marginal_rates (plotRJforward.v3.py:92-139) on already split rows.

    ``rows`` is a list of 1-D float arrays ``[rates..., shifts...]``.  Returns the matrix
    of per-bin rates (oldest bin first, i.e. already reversed back to calendar order).
    """
    nbins = abs(int(end_age - start_age))
    edges = np.arange(end_age, start_age) if end_age < start_age else np.arange(start_age, end_age)
    n = len(rows)
    b = min(int(burnin * n), int(0.9 * n)) if burnin < 1 else int(burnin)
    out = []
    for row in rows[b:]:
        row = np.asarray(row, dtype=np.float64)
        if len(row) == 1:
            out.append(np.zeros(nbins) + row[0])
            continue
        nr = int(np.ceil(len(row) / 2.))
        rates, shifts = row[:nr], row[nr:]
        h = np.histogram(shifts, bins=edges)[0]
        out.append(rates[np.cumsum(h)])
    return np.array(out)


# --------------------------------------------------------------------------------------
# integer accumulators of the one-pass binning (include/literate_b200.h: lr_bin_accumulate /
# lr_bin_finalize) restated with Python integers -- the checker of the lineage-sharded path, where
# shards are combined by an integer SUM before the per-bin values are reconstructed.  The identity is
# SURVEY 7.3; bin_stats() above (the reference's per-bin formulation) stays the ground truth.
# --------------------------------------------------------------------------------------
ACC_ROWS = 8
FIX = 1 << 52


def acc_stride(n_bins):
    return (n_bins + 1 + 7) // 8 * 8


def raw_accumulators(ts, te, first_bin, n_bins, fe_ref=0.5, dead_only=False, end_time=None):
    """int64 [8, acc_stride]: rows births, deaths, sum frac(ts) (low 32 bits / rest, 2^-52 units),
    sum (fe - fe_ref) (same split), births/deaths of lineages without time at risk; [6][n_bins] = lineages
    alive before bin 0."""
    S = acc_stride(n_bins)
    acc = [[0] * S for _ in range(ACC_ROWS)]
    ref_fix = int(round(fe_ref * FIX))
    T0, T1 = first_bin, first_bin + n_bins
    cs = [0] * n_bins
    ce = [0] * n_bins
    for s, e in zip(np.asarray(ts, float).tolist(), np.asarray(te, float).tolist()):
        if dead_only and not (e < end_time):
            continue
        if e > s:
            if s >= T1:
                continue
            if s >= T0:
                a = math.floor(s) - first_bin
                acc[0][a] += 1
                cs[a] += int(round((s - math.floor(s)) * FIX))
            else:
                if not (e > T0):
                    continue
                acc[6][n_bins] += 1
            if e <= T1:
                b = math.ceil(e) - 1 - first_bin
                acc[1][b] += 1
                ce[b] += int(round((e - (math.ceil(e) - 1)) * FIX)) - ref_fix
        else:                                   # no time at risk (te <= ts, or NaN): events still count (:120-121)
            if s >= T0 and s < T1:
                a = math.floor(s) - first_bin
                acc[0][a] += 1
                acc[6][a] += 1
            if e > T0 and e <= T1:
                b = math.ceil(e) - 1 - first_bin
                acc[1][b] += 1
                acc[7][b] += 1
    for j in range(n_bins):
        acc[2][j], acc[3][j] = cs[j] & 0xffffffff, cs[j] >> 32
        acc[4][j], acc[5][j] = ce[j] & 0xffffffff, ce[j] >> 32
    return np.array(acc, dtype=np.int64)


def finalize_accumulators(acc, n_bins, fe_ref=0.5):
    """(sp, ex, br) from (possibly summed) raw accumulators; br is the correctly rounded exact sum."""
    a = [[int(v) for v in row] for row in np.asarray(acc).tolist()]
    ref_fix = int(round(fe_ref * FIX))
    sp = np.array(a[0][:n_bins], dtype=np.int64)
    ex = np.array(a[1][:n_bins], dtype=np.int64)
    br = np.zeros(n_bins)
    c = a[6][n_bins]
    d_prev = 0
    for j in range(n_bins):
        D, E = a[0][j] - a[6][j], a[1][j] - a[7][j]
        c += d_prev - E
        cS = a[2][j] + (a[3][j] << 32)
        cE = a[4][j] + (a[5][j] << 32)
        F = ((c + D) << 52) - cS + E * ref_fix + cE
        br[j] = max(F, 0) / FIX            # int / int: correctly rounded
        d_prev = D
    return sp, ex, br
