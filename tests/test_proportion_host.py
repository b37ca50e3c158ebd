"""CPU: host side of the `-proportion 1` variant (literate_b200/proportion.py) against the oracle restatement that
tests/test_oracle_proportion_golden.py pins to the unmodified LiteRateForward-proportion.py, and against the reference's own
div.log bytes."""
import os

import numpy as np

from conftest import GOLD
from oracle import proportion_oracle as PO
from literate_b200 import proportion as PR

TABLE = os.path.join(GOLD, "proportion", "two_series.tsv")


def test_series_statistics_equal_the_oracle_and_the_reference_div_log(tmp_path):
    ts, te, start, end = PR.read_series(TABLE)
    lin = PO.read_series(TABLE)
    assert (start, end) == (lin.start_time, lin.end_time)
    sp, ex, kn, kd = PR.series_stats(ts, te, start, end)
    st = PO.series_stats(lin)
    for a, b in ((sp, st.sp), (ex, st.ex), (kn, st.kn), (kd, st.kd)):
        assert np.array_equal(a, b, equal_nan=True)
    p = os.path.join(str(tmp_path), "div.log")
    PR.write_div_log(p, sp, ex, kn, kd)
    assert open(p, "rb").read() == open(os.path.join(GOLD, "proportion", "pr_default", "two_series_PR_seed1_div.log"), "rb").read()


def test_series_statistics_on_random_tables_with_gaps(tmp_path):
    rng = np.random.default_rng(8)
    for trial in range(20):
        y0 = int(rng.integers(1900, 2000)); span = int(rng.integers(6, 40))
        n1, n2 = int(rng.integers(5, 200)), int(rng.integers(5, 200))
        s1 = rng.integers(y0, y0 + span, n1); s2 = rng.integers(y0 + rng.integers(0, 3), y0 + span, n2)
        n = max(n1, n2)
        p = os.path.join(str(tmp_path), "t%d.tsv" % trial)
        with open(p, "w") as fh:
            fh.write("id\ta\tb\n")
            for i in range(n):
                fh.write("%d\t%s\t%s\n" % (i, s1[i] if i < n1 else "", s2[i] if i < n2 else ""))
        ts, te, start, end = PR.read_series(p, 0.5)
        got = PR.series_stats(ts, te, start, end)
        st = PO.series_stats(PO.read_series(p, 0.5))
        for a, b in zip(got, (st.sp, st.ex, st.kn, st.kd)):
            assert np.array_equal(a, b, equal_nan=True), trial
        assert len(got[0]) == int(end) - int(start)
        tabs = PR.likelihood_tables(*got)
        assert not np.isnan(tabs["A_birth"]).any() and not np.isnan(tabs["A_death"]).any()     # NaN years are masked out (kn > 0 is False for NaN)
