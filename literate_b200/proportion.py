"""Host side of the `-proportion 1` variant (LiteRateForward-proportion.py): the `ts` and `te` columns are two
independent series of event years, each a piecewise-constant Poisson process of its own.  Only the statistics and the
likelihood tables differ from LiteRateForward.py -- the chains are the same K3 chains on a dataset built from general
tables (engine.Dataset.from_tables).  Plumbing only: the likelihood is evaluated by the CUDA kernels.
"""
from __future__ import annotations

import csv

import numpy as np


def read_series(path, death_jitter=0.5):
    """LiteRateForward-proportion.py:457-492 with -proportion 1: tab-delimited table, empty cells are NaN, te is NOT jittered,
    the window spans both columns and its end moves by the jitter.  Returns (ts, te, start_time, end_time)."""
    t = np.genfromtxt(path, delimiter="\t", skip_header=1)
    if t.ndim == 1:
        t = t[None, :]
    ts, te = (t[:, 2], t[:, 3]) if t.shape[1] == 4 else (t[:, 1], t[:, 2])
    start = float(min(np.nanmin(ts), np.nanmin(te)))
    end = float(max(np.nanmax(ts), np.nanmax(te))) + death_jitter
    return ts, te, start, end


def _interpolated_counts(present, n_years):
    """A year -> count map with None for "year listed but no count": linear interpolation along the sorted years the way
    pandas.Series.interpolate() does it (interior and trailing gaps filled, leading gap left NaN)."""
    keys = sorted(present)
    v = np.array([np.nan if present[k] is None else present[k] for k in keys], dtype=np.float64)
    ok = ~np.isnan(v)
    if ok.any():
        idx = np.arange(len(v))
        filled = np.interp(idx, idx[ok], v[ok])
        filled[:idx[ok][0]] = np.nan
        v = filled
    return v


def series_stats(ts, te, start_time, end_time):
    """:585-598: yearly counts of both series over bins = arange(start, end + 1); a year missing from the second series is a gap
    there, a year missing from the first is 0 when the second series has it and a gap otherwise (the if / elif / if of
    :588-591); gaps interpolated; the last two years dropped.  Returns (sp, ex, kn, kd)."""
    def counts(x):
        u, c = np.unique(x[~np.isnan(x)], return_counts=True)
        return dict(zip(u.tolist(), c.astype(np.float64).tolist()))
    sp, ex = counts(ts), counts(te)
    for y in np.arange(start_time, end_time + 1).tolist():
        if y not in ex:
            ex[y] = None
        elif y not in sp:
            sp[y] = 0.0
        if y not in sp:
            sp[y] = None
    sp_v, ex_v = _interpolated_counts(sp, 0)[:-2], _interpolated_counts(ex, 0)[:-2]
    return sp_v, ex_v, np.cumsum(sp_v), np.cumsum(ex_v)


def likelihood_tables(sp, ex, kn, kd):
    """:157-162 as general tables: events where the running total is positive, exposure Tk = 1 there (model_BDI is forced to 1,
    :442-444: the first series' rate is the immigration rate), no constant; adequacy regresses on the counts (:627-628)."""
    a, b = kn > 0, kd > 0
    return dict(A_birth=np.where(a, sp, 0.0), B_birth=a.astype(np.float64), A_death=np.where(b, ex, 0.0), B_death=b.astype(np.float64),
                x_birth=sp, x_death=ex)


def write_div_log(path, sp, ex, kn, kd):
    """:600-605: four columns; header ends in '\\n', the csv.writer rows in '\\r\\n'."""
    with open(path, "w", newline="") as fh:
        fh.write('sp_events1\tsp_events2\tbr_length1\tbr_length2\n')
        w = csv.writer(fh, delimiter='\t')
        for row in zip(sp.tolist(), ex.tolist(), kn.tolist(), kd.tolist()):
            w.writerow([repr(float(x)) for x in row])
