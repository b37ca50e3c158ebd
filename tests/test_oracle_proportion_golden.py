"""CPU: oracle/proportion_oracle.py (the `-proportion 1` variant) against the log files of the UNMODIFIED
LiteRateForward-proportion.py on the synthetic two-series table (tests/golden/proportion/, oracle/make_golden_proportion.py):
all four files of every run, byte for byte."""
import json
import os

import numpy as np
import pytest

from conftest import GOLD
from oracle import literate_oracle as O
from oracle import proportion_oracle as PO

PR = os.path.join(GOLD, "proportion")


def _cfg(args):
    a = dict(zip(args[::2], args[1::2]))
    return (O.ChainConfig(n_iterations=int(a["-n"]), s_freq=int(a["-s"]), calc_adequacy=int(a.get("-calc_adequacy", 1)),
                          const_death_rate=int(a.get("-const_death_rate", 0)), Poisson_prior=float(a.get("-Poisson_prior", 0)),
                          use_rate_HP=int(a.get("-use_rate_HP", 1)), exact_scipy=True), int(a["-seed"]), float(a.get("-death_jitter", 0.5)))


@pytest.mark.parametrize("run", json.load(open(os.path.join(PR, "manifest.json"))), ids=lambda r: r["tag"])
def test_proportion_oracle_reproduces_the_reference_logs(run, tmp_path):
    cfg, seed, jitter = _cfg(run["args"])
    stem, st, logs = PO.run_reference_style(os.path.join(PR, "two_series.tsv"), str(tmp_path), seed, cfg, death_jitter=jitter)
    for f in run["files"]:
        want = open(os.path.join(PR, run["tag"], f), "rb").read()
        got = open(os.path.join(str(tmp_path), f), "rb").read()
        assert got == want, f


def test_series_statistics():
    """Gap years are interpolated, the last two years dropped, cumulative sums feed the masks (:585-598)."""
    lin = PO.read_series(os.path.join(PR, "two_series.tsv"))
    st = PO.series_stats(lin)
    assert lin.start_time == 1975.0 and lin.end_time == 2019.5 and st.n_bins == 44 == int(lin.end_time) - int(lin.start_time)
    assert np.all(st.kn > 0) and np.all(np.diff(st.kn) >= 0)
    # year 1978 and 1979 (indices 3, 4) have no event in the second series: linear between 1977 and 1980
    assert st.ex[3] == pytest.approx(st.ex[2] + (st.ex[5] - st.ex[2]) / 3) and st.ex[4] == pytest.approx(st.ex[2] + 2 * (st.ex[5] - st.ex[2]) / 3)
