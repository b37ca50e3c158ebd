"""CPU restatement of LiteRateForward-proportion.py with `-proportion 1`.  TEST INFRASTRUCTURE ONLY.

The variant is the sampler of LiteRateForward.py (same proposals, priors, accept rule, log rows: the two files differ only in
the lines cited below) run on OTHER statistics and another likelihood: the `ts` and `te` columns are read as two independent
series of event years ("numerator" and "denominator" immigration processes), each modelled as a piecewise-constant Poisson
process of its own,

    lik = sum_{kn>0} [ log(I_j) U_j - I_j ]  +  sum_{kd>0} [ log(M_j) D_j - M_j ]          (LiteRateForward-proportion.py:157-162)

with U, D the per-year counts of the two columns -- missing years linearly interpolated by pandas, the last two years
dropped (:585-596) -- and kn, kd their running totals (:597-598).  Everything else is oracle.literate_oracle.run_chain.

Pinned: oracle/make_golden_proportion.py runs the UNMODIFIED script on a synthetic two-series table (the tree ships no input
of this kind) and tests/test_oracle_proportion_golden.py requires this module to reproduce its four log files byte for byte.
"""
from __future__ import annotations

import io
import os
from dataclasses import dataclass

import numpy as np

from oracle import literate_oracle as O


@dataclass
class ProportionStats:
    first_bin: int
    sp: np.ndarray      # interpolated yearly counts of the first series (float)
    ex: np.ndarray      # ... of the second series
    kn: np.ndarray      # cumulative sums
    kd: np.ndarray

    @property
    def n_bins(self):
        return len(self.sp)


def read_series(path, death_jitter=0.5):
    """:457-492 with -proportion 1: tab-delimited, empty cells are NaN, te is NOT jittered, the window spans both columns and
    its end is moved by the jitter."""
    t = np.genfromtxt(path, delimiter="\t", skip_header=1)
    if t.shape[1] == 4:
        ts, te = t[:, 2], t[:, 3]
    else:
        ts, te = t[:, 1], t[:, 2]
    start = float(np.min([np.min(ts[~np.isnan(ts)]), np.min(te[~np.isnan(te)])]))
    end = float(np.max([np.max(ts[~np.isnan(ts)]), np.max(te[~np.isnan(te)])])) + death_jitter
    return O.Lineages(ts=ts, te=te, start_time=start, end_time=end, true_root_age=0.0)


def series_stats(lin):
    """:585-598 without pandas: value_counts per year, years of `bins` missing from a series become NaN (0 for the first series
    only when the year is present in the second one and ... exactly as the loop at :588-591 decides), linear interpolation
    along the sorted index (pandas' default: interior and trailing gaps filled, leading NaN kept), last two entries dropped."""
    bins = np.arange(lin.start_time, lin.end_time + 1)

    def counts(x):
        x = x[~np.isnan(x)]
        u, c = np.unique(x, return_counts=True)
        return dict(zip(u.tolist(), c.astype(float).tolist()))

    sp, ex = counts(lin.ts), counts(lin.te)
    for y in bins.tolist():                      # :588-591, statement by statement
        if y not in ex:
            ex[y] = None
        elif y not in sp:
            sp[y] = 0.0
        if y not in sp:
            sp[y] = None

    def interp(d):
        keys = sorted(d)
        v = np.array([np.nan if d[k] is None else d[k] for k in keys], dtype=np.float64)
        ok = ~np.isnan(v)
        if ok.any():
            idx = np.arange(len(v))
            first = idx[ok][0]
            filled = np.interp(idx, idx[ok], v[ok])          # linear inside, last valid value carried forward at the end
            filled[:first] = np.nan                          # pandas leaves leading NaN
            v = filled
        return v[:-2]

    sp_v, ex_v = interp(sp), interp(ex)
    return ProportionStats(int(lin.start_time), sp_v, ex_v, np.cumsum(sp_v), np.cumsum(ex_v))


def loglik(lam_bins, mu_bins, st: ProportionStats):
    """:157-162 (model_BDI forced to 1 at :442-444: I = lambda, Tk = 1)."""
    a, b = st.kn > 0, st.kd > 0
    return np.sum(np.log(lam_bins[a]) * st.sp[a] - lam_bins[a]) + np.sum(np.log(mu_bins[b]) * st.ex[b] - mu_bins[b])


def write_div_log(fh, st: ProportionStats):
    """:600-605: four columns, csv.writer rows (CRLF), floats through str()."""
    import csv
    fh.write("sp_events1\tsp_events2\tbr_length1\tbr_length2\n")
    w = csv.writer(fh, delimiter="\t")
    for row in zip(st.sp, st.ex, st.kn, st.kd):
        w.writerow(row)


def run_reference_style(path, out_dir, seed, cfg: O.ChainConfig, death_jitter=0.5, out=""):
    """The four log files of `LiteRateForward-proportion.py -proportion 1 -seed <seed>`, named as :445, :503-530."""
    lin = read_series(path, death_jitter)
    st = series_stats(lin)
    os.makedirs(out_dir, exist_ok=True)
    stem = os.path.join(out_dir, os.path.splitext(os.path.basename(path))[0] + "_PR_seed" + str(seed) + out)
    with open(stem + "_div.log", "w", newline="") as fh:
        write_div_log(fh, st)
    logs = O.ChainLogs()
    logs.mcmc.write(O.mcmc_header(cfg.calc_adequacy))
    cfg.model_BDI = 1
    shim = O.BinStats(st.first_bin, st.sp, st.ex, st.kn)          # run_chain only takes n_bins from it
    O.run_chain(lin, shim, cfg, seed, logs, lik_fn=lambda l, m: loglik(l, m, st), emp=(st.sp, st.ex))   # :627-628: B_EMP, D_EMP = the counts
    for tag, buf in (("mcmc", logs.mcmc), ("sp_rates", logs.sp), ("ex_rates", logs.ex)):
        with open(stem + "_" + tag + ".log", "w") as fh:
            fh.write(buf.getvalue())
    return stem, st, logs
