"""GPU: the `-proportion 1` variant (LiteRateForward-proportion.py:157-162, :585-604) through the C ABI:
general likelihood tables (lr_dataset_create_general_host) + the unchanged K3 chains.  Oracle: oracle/proportion_oracle.py,
pinned byte for byte to the unmodified script (tests/test_oracle_proportion_golden.py); chain level: 32 unmodified reference
chains (tests/golden/proportion/posterior.json)."""
import json
import os
import shutil

import numpy as np
import pytest

from conftest import GOLD, random_states
from oracle import literate_oracle as O
from oracle import proportion_oracle as PO
from literate_b200 import engine as E, forward as F, proportion as PR

pytestmark = pytest.mark.gpu
TABLE = os.path.join(GOLD, "proportion", "two_series.tsv")


def _dataset(device, path=TABLE, jitter=0.5):
    ts, te, start, end = PR.read_series(path, jitter)
    sp, ex, kn, kd = PR.series_stats(ts, te, start, end)
    ds = E.Dataset.from_tables(device, start_time=start, end_time=end, model_tag=1, **PR.likelihood_tables(sp, ex, kn, kd))
    return ds, PO.series_stats(PO.read_series(path, jitter)), start, end


def test_likelihood_and_adequacy_match_the_oracle(device):
    """Random states: likelihood within 1e-10 relative of :157-162 as the oracle restates it; adequacy (regression on the
    COUNTS, :627-628) within 1e-10 of the closed form of calculate_r_squared."""
    ds, st, start, end = _dataset(device)
    rng = np.random.default_rng(3)
    states = random_states(rng, 200, start, end, kmax=9, rate_scale=6.0)
    got = ds.evaluate(states)
    for i, (L, M, tL, tM) in enumerate(states):
        iL = O.rate_index(np.floor(tL) if len(tL) > 2 else tL, st.n_bins)
        iM = O.rate_index(np.floor(tM) if len(tM) > 2 else tM, st.n_bins)
        assert got["lik"][i] == pytest.approx(PO.loglik(L[iL], M[iM], st), rel=1e-10)
        np.testing.assert_allclose(got["adequacy"][i], O.adequacy_closed_form(st.sp, st.ex, L[iL], M[iM]), rtol=1e-10)


def test_late_starting_series_is_masked(device, tmp_path):
    p = os.path.join(str(tmp_path), "late.tsv")
    rng = np.random.default_rng(5)
    s1 = rng.integers(1980, 2010, 300); s2 = rng.integers(1986, 2010, 120)
    with open(p, "w") as fh:
        fh.write("id\ta\tb\n")
        for i in range(300):
            fh.write("%d\t%d\t%s\n" % (i, s1[i], s2[i] if i < 120 else ""))
    ds, st, start, end = _dataset(device, p)
    assert np.isnan(st.ex[:3]).all() and np.isnan(st.kd[:3]).all()
    states = random_states(rng, 50, start, end, kmax=5, rate_scale=5.0)
    got = ds.evaluate(states)
    for i, (L, M, tL, tM) in enumerate(states):
        iL = O.rate_index(np.floor(tL) if len(tL) > 2 else tL, st.n_bins)
        iM = O.rate_index(np.floor(tM) if len(tM) > 2 else tM, st.n_bins)
        assert np.isfinite(got["lik"][i]) and got["lik"][i] == pytest.approx(PO.loglik(L[iL], M[iM], st), rel=1e-10)


def test_command_line_writes_the_reference_layout(device, tmp_path):
    """`python -m literate_b200.forward -d table -proportion 1`: files <table>_PR_seed<seed>_{mcmc,sp_rates,ex_rates,div}.log, div.log
    byte-identical to the unmodified script's, mcmc.log with its header and one row per sample, every logged state consistent."""
    src = os.path.join(str(tmp_path), "two_series.tsv")
    shutil.copy(TABLE, src)
    args = F.build_parser().parse_args(["-d", src, "-proportion", "1", "-n", "4001", "-s", "100", "-p", "100000", "-seed", "3", "-quiet", "1"])
    paths = F.run(args, device)
    d = os.path.join(str(tmp_path), "literate_mcmc_logs")
    assert sorted(os.listdir(d)) == ["two_series_PR_seed3_%s.log" % t for t in ("div", "ex_rates", "mcmc", "sp_rates")]
    ref = os.path.join(GOLD, "proportion", "pr_default")
    assert open(os.path.join(d, "two_series_PR_seed3_div.log"), "rb").read() == open(os.path.join(ref, "two_series_PR_seed1_div.log"), "rb").read()
    mine = open(paths[0]).read().splitlines()
    want = open(os.path.join(ref, "two_series_PR_seed1_mcmc.log")).read().splitlines()
    assert mine[0] == want[0] and len(mine) == 1 + 41
    st = PO.series_stats(PO.read_series(TABLE))
    sp_rows = open(paths[0].replace("mcmc.log", "sp_rates.log")).read().splitlines()
    ex_rows = open(paths[0].replace("mcmc.log", "ex_rates.log")).read().splitlines()
    for row, a, b in list(zip(mine[1:], sp_rows, ex_rows))[5:]:
        f = [float(x) for x in row.split("\t")]
        K_l, K_m = int(f[6]), int(f[7])
        a = np.array(a.split("\t"), float); b = np.array(b.split("\t"), float)
        L, tL = a[:K_l], np.concatenate([[f[8]], a[K_l:], [f[9]]])
        M, tM = b[:K_m], np.concatenate([[f[8]], b[K_m:], [f[9]]])
        iL = O.rate_index(np.floor(tL) if K_l > 1 else tL, st.n_bins)
        iM = O.rate_index(np.floor(tM) if K_m > 1 else tM, st.n_bins)
        assert f[2] == pytest.approx(PO.loglik(L[iL], M[iM], st), rel=1e-10) and f[1] == pytest.approx(f[2] + f[3], rel=1e-12)


def test_posterior_matches_the_reference_chains(device):
    """128 device chains against 32 chains of the unmodified script (200 001 iterations, 20 % burn-in) on the two-series table:
    the comparison of tests/test_gpu_chains.py (means, per-bin marginal rates, K distributions, likelihood variance, 3.5 s.e.)."""
    from test_gpu_chains import _compare_with_reference, _summaries
    p = os.path.join(GOLD, "proportion", "posterior.json")
    if not os.path.exists(p):
        pytest.skip("fixture not generated (oracle/make_golden_proportion.py posterior)")
    ref = json.load(open(p))["chains"]
    ds, st, start, end = _dataset(device)
    ch = E.Chains(ds, 128, seed=515, cfg=E.default_config(1))
    recs = ch.run(ref[0]["n_iterations"], ref[0]["s_freq"])
    lin = O.Lineages(ts=None, te=None, start_time=start, end_time=end, true_root_age=0.0)
    _compare_with_reference(_summaries(recs, lin), ref)
