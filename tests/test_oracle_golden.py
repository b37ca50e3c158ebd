"""CPU: the oracle against everything the reference offers for this path.

* the worked example of tutorial 2 (2_introduction_to_literate_final.ipynb:175-216) -- the only
  known-answer vector the reference contains;
* the log files the UNMODIFIED reference wrote in the build container (tests/golden/reference_logs,
  made by oracle/make_golden.py): the oracle chain, with the same seed, must reproduce them byte
  for byte (sufficient statistics, likelihoods, priors, proposals, accept rule, log formatting).
"""
import io
import json
import os

import numpy as np
import pytest

from conftest import GOLD, golden_input
from oracle import literate_oracle as O

# 100 (birth, death) times of the tutorial cell; frame [10, 40], lambda=.12, mu=.04 -> B=42, D=9, S=462.68
NOTEBOOK_S = 462.68
NOTEBOOK_LOGLIK = -192.04975094421764


def _flag(args, name, default, cast=float):
    return cast(args[args.index(name) + 1]) if name in args else default


def _jobs():
    with open(os.path.join(GOLD, "reference_logs", "manifest.json")) as fh:
        return json.load(fh)


@pytest.mark.parametrize("job", _jobs(), ids=lambda j: j["tag"])
def test_oracle_chain_reproduces_reference_logs(job, tmp_path):
    a = job["args"]
    cfg = O.ChainConfig(n_iterations=_flag(a, "-n", 10000000, int), s_freq=_flag(a, "-s", 1000, int),
                        model_BDI=_flag(a, "-model_BDI", 0, int), const_rates=_flag(a, "-const_rates", 0, int),
                        const_death_rate=_flag(a, "-const_death_rate", 0, int), use_rate_HP=_flag(a, "-use_rate_HP", 1, int),
                        Poisson_prior=_flag(a, "-Poisson_prior", 0.0), calc_adequacy=_flag(a, "-calc_adequacy", 1, int),
                        update_fraction=_flag(a, "-update_fraction", 0.75), pyrate_output="-pyrate_output" in a,
                        exact_scipy=True)
    src = golden_input(job["input"], tmp_path)
    out_dir = os.path.join(str(tmp_path), "out")
    O.run_reference_style(src, out_dir, _flag(a, "-seed", -1, int), cfg, TBP="-TBP" in a,
                          death_jitter=_flag(a, "-death_jitter", 0.5), out=_flag(a, "-out", "", str))
    for f in job["files"]:
        want = open(os.path.join(GOLD, "reference_logs", job["tag"], f), "rb").read()
        got = open(os.path.join(out_dir, f), "rb").read()
        assert got == want, f"{job['tag']}/{f} differs from the reference's own output"


def test_closed_form_densities_match_scipy():
    import scipy.stats
    rng = np.random.default_rng(0)
    x = rng.gamma(2.0, 1.0, 50) + 1e-3
    for rate in (0.3, 1.0, 2.0, 17.5):
        np.testing.assert_allclose(O.ln_gamma_pdf(x, 2.0, rate), scipy.stats.gamma.logpdf(x, 2.0, scale=1. / rate), rtol=5e-14, atol=5e-14)
    for u in rng.uniform(0.01, 0.99, 50):
        assert abs(O.ln_sym_beta_pdf(u, 10.0) - scipy.stats.beta.logpdf(u, 10.0, 10.0)) < 5e-13


def test_example_tad_sufficient_statistics():
    lin = O.read_lineages(golden_input("example_dataTAD.txt"))
    st = O.bin_stats(lin.ts, lin.te)
    assert st.first_bin == 1994 and st.n_bins == 24
    assert st.sp.sum() == 75 and st.ex.sum() == 61 and st.br.sum() == 361.5
    assert list(zip(st.sp[:3], st.ex[:3], st.br[:3])) == [(2, 0, 2.0), (1, 0, 3.0), (8, 5, 8.5)]
    fast = O.bin_stats_fast(lin.ts, lin.te)
    assert (fast.sp == st.sp).all() and (fast.ex == st.ex).all() and (fast.br == st.br).all()


def test_metal_bands_totals(metal_path):
    lin = O.read_lineages(metal_path)
    st = O.bin_stats(lin.ts, lin.te, only_dead=True, end_time=lin.end_time)
    assert st.first_bin == 1968 and st.n_bins == 32
    assert st.sp.sum() == 27495 and st.ex.sum() == 16191 and st.br.sum() == 95426.5
    fast = O.bin_stats_fast(lin.ts, lin.te, only_dead=True, end_time=lin.end_time)
    for a, b in ((fast.sp, st.sp), (fast.ex, st.ex), (fast.br, st.br), (fast.ex_dead, st.ex_dead), (fast.br_dead, st.br_dead)):
        assert (a == b).all()


def test_initial_state_likelihoods_of_survey_8c():
    lin = O.read_lineages(golden_input("example_dataTAD.txt"))
    st = O.bin_stats(lin.ts, lin.te, only_dead=True, end_time=lin.end_time)
    L, M = np.array([9.255751002593213]), np.array([1.9901459055735302])
    t = np.array([lin.start_time, lin.end_time])
    want = {0: -3488.0146007277763, 1: -563.4000768986649, 2: -3856.517665395376, 3: -3778.901975078008}
    for m, v in want.items():
        assert O.loglik_state(L, M, t, t, st, m) == pytest.approx(v, rel=1e-14)
    # the `prior` column of row 0 of the reference's log: Gamma_rate = [1, 1] (:222, :296), not the rate 2 of :227
    prior = O.state_prior(L, M, [1., 1.], lin.end_time - lin.start_time, 2 * O.poisson_prior(1, 1))
    assert prior == pytest.approx(-10.332443864394012, rel=1e-14)


def test_notebook_known_answer():
    """Frame likelihood of tutorial 2: B log(l) + D log(m) - (l+m) S with B=42, D=9, S=462.68."""
    ll = 42 * np.log(.12) + 9 * np.log(.04) - (.12 + .04) * NOTEBOOK_S
    assert ll == pytest.approx(NOTEBOOK_LOGLIK, rel=1e-13)
    assert np.exp(ll) == pytest.approx(3.9251197773152857e-84, rel=1e-11)
    # the same number through the oracle's Keiding likelihood on a single bin carrying those statistics
    st = O.BinStats(10, np.array([42]), np.array([9]), np.array([NOTEBOOK_S]))
    assert O.loglik_keiding(np.array([.12]), np.array([.04]), st, False) == pytest.approx(NOTEBOOK_LOGLIK, rel=1e-13)


def test_fast_binning_equals_loop_on_random_data():
    rng = np.random.default_rng(5)
    for trial in range(6):
        n = int(rng.integers(1, 400))
        ts = 100 + np.floor(rng.uniform(0, 40, n))
        te = np.minimum(ts + np.floor(rng.exponential(6, n)), 140) + (0.5 if trial % 2 == 0 else 0.0)
        if trial == 5:
            ts = ts + rng.integers(0, 4, n) / 4.0          # dyadic fractions: still exact
            te = np.maximum(te, ts)
        a, b = O.bin_stats(ts, te, only_dead=True), O.bin_stats_fast(ts, te, only_dead=True)
        for x, y in ((a.sp, b.sp), (a.ex, b.ex), (a.br, b.br), (a.ex_dead, b.ex_dead), (a.br_dead, b.br_dead)):
            assert (x == y).all()


def test_adequacy_closed_form_matches_lstsq():
    rng = np.random.default_rng(9)
    for _ in range(5):
        eb, ed, sb, sd = (rng.gamma(2, 1, 24) for _ in range(4))
        np.testing.assert_allclose(O.adequacy_closed_form(eb, ed, sb, sd), O.adequacy(eb, ed, sb, sd), rtol=1e-10)
