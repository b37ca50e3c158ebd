"""CPU oracle for TrendRate: fixed-dimension Metropolis-Hastings on the binned statistics with birth and
death rates that follow an exogenous trend.  TEST INFRASTRUCTURE ONLY.

NumPy restatement of ``trend_rate.py`` and of the ``literate_library.py`` functions it calls.  The script cannot be
imported (it parses ``sys.argv`` and runs the chain at module level).  Only ``tests/`` and ``__graft_entry__.smoke()`` may import this
module, as the checker -- never the product path.

Parity pin: ``oracle/make_golden_trend.py`` runs the *unmodified* ``trend_rate.py`` in the build container (the tree
ships no trend table, so the script writes a small synthetic one next to a three-column copy of the example table) and
``run_chain`` below, with the same seed and the same legacy ``np.random`` draw order, reproduces its log byte for byte;
fixtures under ``tests/golden/trendrate/``, re-checked by ``tests/test_oracle_trend_golden.py``.

All ``file:line`` citations are relative to the reference checkout.
"""
from __future__ import annotations

import csv
import math
import os
from dataclasses import dataclass

import numpy as np

SMALL_NUMBER = 0.000000000000001     # trend_rate.py:55
PARAMS = ["l_min", "m_min", "alpha", "beta", "delta", "gamma"]        # trend_rate.py:74
LN_2PI = math.log(2.0 * math.pi)


# --------------------------------------------------------------------------------------
# input (literate_library.py:196-229, :231-257; trend_rate.py:58-69)
# --------------------------------------------------------------------------------------
def parse_ts_te(path, TBP=False, first_year=-1, last_year=-1, death_jitter=0.5):
    """literate_library.py:196-229 (pandas.read_csv with a tab delimiter: the first line is the header)."""
    import pandas as pd         # the reference's parser: its fast float conversion is not always correctly rounded
    t = pd.read_csv(path, delimiter="\t").to_numpy()
    if t.shape[1] == 4:                 # :198-201
        ts_y, te_y = t[:, 2], t[:, 3]
    else:                               # :203-204
        ts_y, te_y = t[:, 1], t[:, 2]
    if TBP:                             # :205-214
        if first_year != -1:
            te_y = te_y[ts_y <= first_year]
            ts_y = ts_y[ts_y <= first_year]
        if last_year != -1:             # (the reference masks te with the already filtered ts here, :210-211)
            keep = ts_y >= last_year
            ts_y, te_y = ts_y[keep], te_y[keep].copy()
            te_y[te_y < last_year] = last_year
        root = np.max(ts_y)
        ts, te = root - ts_y, root - te_y
    else:                               # :216-224
        if first_year != -1:
            te_y = te_y[ts_y >= first_year]
            ts_y = ts_y[ts_y >= first_year]
        if last_year != -1:
            te_y = te_y[ts_y <= last_year].copy()
            ts_y = ts_y[ts_y <= last_year]
            te_y[te_y > last_year] = last_year
        ts, te = ts_y, te_y
    te = te + death_jitter              # :226
    return ts, te, float(np.max(te)), float(np.min(ts))     # present, origin (:227-229)


def precompute_events(ts, te, t0, t1):
    """literate_library.py:74-85."""
    sp = int(np.count_nonzero((ts >= t0) & (ts < t1)))
    ex = int(np.count_nonzero((te > t0) & (te <= t1)))
    s = np.maximum(ts, t0)
    e = np.minimum(te, t1)
    dt = e - s
    return sp, ex, float(np.sum(dt[dt > 0]))


@dataclass
class Bins:
    origin: float
    present: float
    n_spec: np.ndarray      # int64 [n_bins]
    n_exti: np.ndarray
    dt: np.ndarray          # fp64 [n_bins]

    @property
    def n_bins(self):
        return len(self.dt)


def create_bins(origin, present, ts, te, rm_first_bin=0):
    """literate_library.py:231-257: unit bins from ``origin``; the last one is always dropped."""
    edges = np.arange(origin, present + 1)
    sp, ex, br = [], [], []
    for i in range(len(edges) - 1):
        a, b, c = precompute_events(ts, te, edges[i], edges[i + 1])
        sp.append(a); ex.append(b); br.append(c)
    sp = np.array(sp, dtype=np.int64)[:-1]
    ex = np.array(ex, dtype=np.int64)[:-1]
    br = np.array(br, dtype=float)[:-1]
    if rm_first_bin:
        sp, ex, br = sp[1:], ex[1:], br[1:]
        origin += 1
    return Bins(origin, present, sp, ex, br)


def read_trend_column(path, index):
    """Column ``index`` of a tab-separated table with a header line (trend_rate.py:59-60)."""
    import pandas as pd         # pandas' default float conversion can differ from float() in the last bit: use the same
    return pd.read_csv(path, sep="\t").iloc[:, index].to_numpy()


def normalise_trend(trend, rm_first_bin=0):
    """trend_rate.py:61-69: drop the last bin, min-max scale, zeros become SMALL_NUMBER."""
    trend = np.asarray(trend, dtype=float)[:-1]
    if rm_first_bin:
        trend = trend[1:]
    lo, hi = np.min(trend), np.max(trend)
    trend = (trend - lo) / (hi - lo)
    trend[trend == 0] = SMALL_NUMBER
    return trend


# --------------------------------------------------------------------------------------
# likelihood and priors (trend_rate.py:73-100)
# --------------------------------------------------------------------------------------
def rates_of(params, trend, const_birth=False, const_death=False):
    l_min, m_min, alpha, beta, delta, gamma = params
    n = len(trend)
    if const_birth:                                   # :75-76
        lam = np.ones(n) * l_min
    else:                                             # :79-80
        lam = l_min + alpha * trend ** delta
        lam[lam <= 0.0] = SMALL_NUMBER
    if const_death:                                   # :83-84
        mu = np.ones(n) * m_min
    else:                                             # :86-87
        mu = m_min + beta * trend ** gamma
        mu[mu <= 0.0] = SMALL_NUMBER
    return lam, mu


def likelihood(params, bins: Bins, trend, const_birth=False, const_death=False):
    """trend_rate.py:73-91 -> ((birth_lik, death_lik), birth_rates, death_rates)."""
    lam, mu = rates_of(params, trend, const_birth, const_death)
    with np.errstate(divide="ignore", invalid="ignore"):
        bl = np.sum(np.log(lam) * bins.n_spec - lam * bins.dt)      # :82
        dl = np.sum(np.log(mu) * bins.n_exti - mu * bins.dt)        # :88
    return np.array([bl, dl]), lam, mu


def ln_gamma_pdf_loc(x, a, scale, loc):
    """scipy.stats.gamma.logpdf(x, a, scale=scale, loc=loc) (literate_library.py:182-184), closed form."""
    y = (x - loc) / scale
    if y < 0 or (y == 0 and a > 1):
        return -math.inf
    if y == 0:
        return (-math.log(scale) - math.lgamma(a)) if a == 1 else math.inf
    return (a - 1.0) * math.log(y) - y - math.lgamma(a) - math.log(scale)


def ln_norm_pdf(x, loc, scale):
    """scipy.stats.norm.logpdf (literate_library.py:186-187)."""
    z = (x - loc) / scale
    return -0.5 * z * z - 0.5 * LN_2PI - math.log(scale)


def prior(params, exact_scipy=False):
    """trend_rate.py:93-100."""
    if exact_scipy:
        import scipy.stats
        p = scipy.stats.gamma.logpdf(params[0], 1, scale=10, loc=.001)
        p += scipy.stats.gamma.logpdf(params[1], 1, scale=10, loc=.001)
        p += scipy.stats.norm.logpdf(params[2], loc=0, scale=5)
        p += scipy.stats.norm.logpdf(params[3], loc=0, scale=5)
        p += scipy.stats.gamma.logpdf(params[4], 3, scale=.5, loc=0)
        p += scipy.stats.gamma.logpdf(params[5], 3, scale=.5, loc=0)
        return p
    p = ln_gamma_pdf_loc(params[0], 1, 10, .001)
    p += ln_gamma_pdf_loc(params[1], 1, 10, .001)
    p += ln_norm_pdf(params[2], 0, 5)
    p += ln_norm_pdf(params[3], 0, 5)
    p += ln_gamma_pdf_loc(params[4], 3, .5, 0)
    p += ln_gamma_pdf_loc(params[5], 3, .5, 0)
    return p


def adequacy(emp_birth, emp_death, est_birth, est_death):
    """calculate_r_squared (literate_library.py:268-279)."""
    x = np.concatenate([emp_birth, emp_death])
    y = np.concatenate([est_birth, est_death])
    res = np.linalg.lstsq(np.vstack(x), y, rcond=None)
    coeff = res[0][0]
    ssres = res[1][0]
    r2 = 1 - ssres / np.sum(y ** 2)
    fitted = coeff * x
    resid = y - fitted
    vf = np.var(fitted, ddof=1)
    return coeff, r2, vf / (vf + np.var(resid, ddof=1))


# --------------------------------------------------------------------------------------
# proposals (literate_library.py:140-146, :156-165) with explicit draws, for deterministic parity of the device arithmetic
# --------------------------------------------------------------------------------------
def move_weights(const_birth=False, const_death=False):
    """trend_rate.py:122-134 -> (update_multiplier, update_normal) per-parameter probabilities."""
    mult = np.array([1, 1, 0, 0, 1, 1], dtype=float)
    if const_birth:
        mult[4] = 0
    if const_death:
        mult[5] = 0
    mult = mult / np.sum(mult)
    norm = np.array([0, 0, 1, 1, 0, 0], dtype=float)
    if const_birth:
        norm[2] = 0
    if const_death:
        norm[3] = 0
    with np.errstate(invalid="ignore", divide="ignore"):
        norm = norm / np.sum(norm)       # 0/0 with both sides constant: the reference then dies in np.random.binomial
    return mult, norm


def multiplier_given(q, on, u, d=1.1):
    """update_multiplier_proposal_vec with given mask and uniforms (literate_library.py:156-165)."""
    m = np.exp(2 * np.log(d) * (np.asarray(u) - .5))
    m[np.asarray(on) == 0] = 1.
    return np.asarray(q) * m, float(np.sum(np.log(m)))


def normal_given(q, on, z, d=.001):
    """update_normal_nobound_vec with given mask and standard normals (literate_library.py:140-146)."""
    m = d * np.asarray(z, dtype=float)
    m[np.asarray(on) == 0] = 0.
    return np.asarray(q) + m, 0.0


# --------------------------------------------------------------------------------------
# the chain (trend_rate.py:102-196), legacy global np.random stream, same draw order
# --------------------------------------------------------------------------------------
def log_name(path, seed, trend_index, const_birth=False, const_death=False, no_death=False):
    out = "_CONB" if const_birth else "_EXPB"            # :104-108
    if no_death:
        out += "_ND"
    elif const_death:
        out += "_COND"
    else:
        out += "_EXPD"
    return "%s_%s%s_%s.trendrate.log" % (os.path.splitext(path)[0], seed, out, trend_index)      # :110


def header(n_bins):
    head = ["it", "posterior", "likelihood", "likelihood_birth", "likelihood_death", "prior"] + PARAMS      # :113
    head += ["l_%s" % i for i in range(n_bins)] + ["m_%s" % i for i in range(n_bins)]
    return head + ["corr_coeff", "rsquared", "gelman_r2"]


def run_chain(bins: Bins, trend, n_iter, sample_every, seed, fh=None, const_birth=False, no_death=False,
              const_death=False, exact_scipy=True, collect=False):
    """The reference chain.  Writes the log rows to ``fh`` (text file opened with newline='') like :111-118, :184-194.

    ``exact_scipy`` evaluates the prior with the reference's scipy calls (bit-equal logs); the closed forms agree to 1e-15.
    Returns the sampled rows when ``collect``."""
    if no_death:
        const_death = True                               # :44
    np.random.seed(seed)                                  # set_seed, literate_library.py:282-288
    with np.errstate(divide="ignore", invalid="ignore"):
        emp_b, emp_d = bins.n_spec / bins.dt, bins.n_exti / bins.dt      # :53, literate_library.py:260-265
    wlog = None
    if fh is not None:
        wlog = csv.writer(fh, delimiter="\t")
        wlog.writerow(header(bins.n_bins))
    f_mult, f_norm = move_weights(const_birth, const_death)
    argsA = np.array([.1, .1, 0, 0, 1, 1], dtype=float)                  # :141-147
    lk, lam, mu = likelihood(argsA, bins, trend, const_birth, const_death)
    likA, likB, likD = np.sum(lk), lk[0], lk[1]
    priorA = prior(argsA, exact_scipy)
    rows = []
    it = 0
    while it != n_iter:                                                   # :161
        args = argsA + 0.
        rr = np.random.random(1)
        if rr[0] < .33:                                                   # :167-168, literate_library.py:140-146
            S = np.shape(args)
            ff = np.random.binomial(1, f_norm, S)
            m = np.random.normal(0, .001, S)
            m[ff == 0] = 0.
            args, hast = args + m, 0
        else:                                                             # :170, literate_library.py:156-165
            S = np.shape(args)
            ff = np.random.binomial(1, f_mult, S)
            u = np.random.uniform(0, 1, S)
            m = np.exp(2 * np.log(1.1) * (u - .5))
            m[ff == 0] = 1.
            args, hast = args * m, np.sum(np.log(m))
        lk2, lam2, mu2 = likelihood(args, bins, trend, const_birth, const_death)
        lik = np.sum(lk2)
        pr = prior(args, exact_scipy)
        if (lik - likA) + (pr - priorA) + hast > np.log(np.random.random()) or it == 0:      # :176
            argsA, priorA, likA, likB, likD, lam, mu = args, pr, lik, lk2[0], lk2[1], lam2, mu2
        if it % sample_every == 0:                                        # :184-194
            with np.errstate(divide="ignore", invalid="ignore"):
                adq = adequacy(emp_b, emp_d, lam, mu)
            row = [it, likA + priorA, likA, likB, likD, priorA] + list(argsA) + list(lam) + list(mu) + list(adq)
            if wlog is not None:
                wlog.writerow(row)
            if collect:
                rows.append(np.array(row, dtype=float))
        it += 1
    return np.array(rows) if collect else None
