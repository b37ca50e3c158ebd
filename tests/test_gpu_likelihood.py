"""GPU: K2 (likelihood + priors on explicit states) through the C ABI against the oracle.
Tolerance: 1e-10 relative (BASELINE.json north_star) -- in practice ~1e-14."""
import numpy as np
import pytest

from conftest import golden_input, random_states
from oracle import literate_oracle as O
from literate_b200.engine import Dataset, BinStats

pytestmark = pytest.mark.gpu
RTOL = 1e-10


def _dataset(device, lin, model):
    st = O.bin_stats(lin.ts, lin.te, only_dead=True, end_time=lin.end_time)
    bs = BinStats(st.first_bin, st.sp[None], st.ex[None], st.br[None], st.ex_dead[None], st.br_dead[None])
    return st, Dataset(device, bs, model, lin.start_time, lin.end_time)


@pytest.mark.parametrize("model", [0, 1, 2, 3])
def test_initial_state_goldens(device, model):
    lin = O.read_lineages(golden_input("example_dataTAD.txt"))
    st, ds = _dataset(device, lin, model)
    t = np.array([lin.start_time, lin.end_time])
    out = ds.evaluate([(np.array([9.255751002593213]), np.array([1.9901459055735302]), t, t)], gamma_rate=[1., 1.], poi_lambda=1.0)
    want = {0: -3488.0146007277763, 1: -563.4000768986649, 2: -3856.517665395376, 3: -3778.901975078008}[model]
    assert out["lik"][0] == pytest.approx(want, rel=RTOL)
    assert out["prior_rates"][0] + out["prior_poi"][0] == pytest.approx(-10.332443864394012, rel=RTOL)


@pytest.mark.parametrize("model", [0, 1, 2, 3])
@pytest.mark.parametrize("data", ["tad", "tbp", "metal"])
def test_random_states(device, model, data, metal_path):
    if data == "tad":
        lin = O.read_lineages(golden_input("example_dataTAD.txt"))
    elif data == "tbp":
        lin = O.read_lineages(golden_input("example_dataTBP.txt"), TBP=True)
    else:
        lin = O.read_lineages(metal_path)
    st, ds = _dataset(device, lin, model)
    rng = np.random.default_rng(100 + model)
    states = random_states(rng, 64, lin.start_time, lin.end_time, kmax=9, rate_scale=0.3)
    g = rng.gamma(2.0, 1.0, (64, 2)) + 0.05
    p = rng.gamma(2.0, 1.0, 64) + 0.05
    out = ds.evaluate(states, gamma_rate=g, poi_lambda=p)
    emp_b, emp_d = st.sp / st.br, st.ex / st.br
    for i, (L, M, tL, tM) in enumerate(states):
        lik = O.loglik_state(L, M, tL, tM, st, model)
        assert out["lik"][i] == pytest.approx(lik, rel=RTOL), (i, len(L), len(M))
        pr = O.state_prior(L, M, g[i], lin.end_time - lin.start_time, 0.0)
        assert out["prior_rates"][i] == pytest.approx(pr, rel=RTOL, abs=1e-12)
        pp = O.poisson_prior(len(L), p[i]) + O.poisson_prior(len(M), p[i])
        assert out["prior_poi"][i] == pytest.approx(pp, rel=RTOL, abs=1e-12)
        iL = O.rate_index(np.floor(tL) if len(tL) > 2 else tL, st.n_bins)
        iM = O.rate_index(np.floor(tM) if len(tM) > 2 else tM, st.n_bins)
        adq = O.adequacy(emp_b, emp_d, L[iL], M[iM])
        np.testing.assert_allclose(out["adequacy"][i], adq, rtol=1e-8)


def test_scipy_priors_agree(device):
    """The same prior through scipy.stats (what the reference calls, LiteRateForward.py:201-202)."""
    import scipy.stats
    lin = O.read_lineages(golden_input("example_dataTAD.txt"))
    st, ds = _dataset(device, lin, 0)
    rng = np.random.default_rng(7)
    states = random_states(rng, 16, lin.start_time, lin.end_time, kmax=6)
    out = ds.evaluate(states, gamma_rate=[0.7, 3.1], poi_lambda=2.0)
    for i, (L, M, tL, tM) in enumerate(states):
        pr = np.sum(scipy.stats.gamma.logpdf(L, 2.0, scale=1 / 0.7)) + np.sum(scipy.stats.gamma.logpdf(M, 2.0, scale=1 / 3.1))
        pr += -np.log(lin.end_time - lin.start_time) * (len(L) + len(M) - 2)
        assert out["prior_rates"][i] == pytest.approx(pr, rel=RTOL)


def test_bins_with_zero_time_at_risk_are_masked(device):
    """BDI_partial_lik drops bins with br == 0 (:160); the Keiding form keeps them."""
    sp = np.array([[3, 0, 2, 0, 1, 4]]); ex = np.array([[1, 0, 1, 0, 0, 2]]); br = np.array([[4.5, 0.0, 3.0, 0.0, 2.5, 6.0]])
    st = O.BinStats(10, sp[0], ex[0], br[0], ex[0], br[0])
    rng = np.random.default_rng(1)
    states = random_states(rng, 20, 10.0, 16.5, kmax=3)
    for model in range(4):
        ds = Dataset(device, BinStats(10, sp, ex, br, ex, br), model, 10.0, 16.5)
        out = ds.evaluate(states)
        for i, (L, M, tL, tM) in enumerate(states):
            assert out["lik"][i] == pytest.approx(O.loglik_state(L, M, tL, tM, st, model), rel=RTOL)


def test_replicates_select_their_own_tables(device):
    rng = np.random.default_rng(2)
    sp = rng.integers(0, 50, (3, 12)); ex = rng.integers(0, 40, (3, 12)); br = rng.uniform(5, 80, (3, 12))
    ds = Dataset(device, BinStats(0, sp, ex, br), 0, 0.0, 12.5)
    states = random_states(rng, 9, 0.0, 12.5, kmax=4)
    rep = np.arange(9) % 3
    out = ds.evaluate(states, rep=rep)
    for i, (L, M, tL, tM) in enumerate(states):
        st = O.BinStats(0, sp[rep[i]], ex[rep[i]], br[rep[i]])
        assert out["lik"][i] == pytest.approx(O.loglik_state(L, M, tL, tM, st, 0), rel=RTOL)


@pytest.mark.parametrize("model", [0, 2])
def test_rj_proposals_match_the_oracle(device, model, metal_path):
    """add_shift_RJ_weighted_mean / remove_shift_RJ_weighted_mean (LiteRateForward.py:29-69) on explicit states with explicit
    draws: proposed rates and times, the Hastings + Jacobian term and the whole acceptance ratio of :296-313 against the
    oracle's restatement (the functions its byte-for-byte pinned chain calls); add followed by the matching remove returns
    to the start with opposite log-ratios."""
    from literate_b200.engine import evaluate_proposals
    lin = O.read_lineages(metal_path)
    st, ds = _dataset(device, lin, model)
    rng = np.random.default_rng(40 + model)
    span = lin.end_time - lin.start_time
    states = random_states(rng, 96, lin.start_time, lin.end_time, kmax=8, rate_scale=0.3)
    n = len(states)
    side = rng.integers(0, 2, n)
    g = rng.gamma(2.0, 1.0, (n, 2)) + 0.05
    lam = rng.gamma(2.0, 1.0, n) + 0.05
    poiA = rng.normal(-3, 1, n)
    Ks = np.array([len(s[0]) if sd else len(s[1]) for s, sd in zip(states, side)])
    kind = np.where((rng.random(n) < 0.5) & (Ks > 1), 3, 2)
    idx = np.array([rng.integers(0, k) if kd == 2 else rng.integers(1, k) for k, kd in zip(Ks, kind)])
    u_t, u_b = rng.uniform(0.02, 0.98, n), rng.beta(10, 10, n)
    out = evaluate_proposals(ds, states, side, kind, idx, u_t, u_b, gamma_rate=g, poi_lambda=lam, poiA=poiA)

    def lik_prior(L, M, tL, tM, i, poi_term):
        return O.loglik_state(L, M, tL, tM, st, model), O.state_prior(L, M, g[i], span, poi_term)

    n_checked = 0
    for i, (L, M, tL, tM) in enumerate(states):
        r, t = (L, tL) if side[i] else (M, tM)
        if kind[i] == 2:
            t_new = t[idx[i]] + u_t[i] * (t[idx[i] + 1] - t[idx[i]])
            guard = min(t_new - t[idx[i]], t[idx[i] + 1] - t_new) <= 1
            assert bool(out["ok"][i]) == (not guard), i
            if guard:
                continue
            rn, tn, h = O.add_shift_given(r, t, idx[i], t_new, u_b[i])
        else:
            rn, tn, h = O.remove_shift_given(r, t, idx[i])
            assert out["ok"][i]
        np.testing.assert_allclose(out["rates"][i], rn, rtol=1e-12)
        np.testing.assert_allclose(out["times"][i], tn, rtol=1e-14)
        assert out["hasting"][i] == pytest.approx(h, rel=1e-11, abs=1e-12)
        L2, M2, tL2, tM2 = (rn, M, tn, tM) if side[i] else (L, rn, tL, tn)
        poiN = O.poisson_prior(len(L2), lam[i]) + O.poisson_prior(len(M2), lam[i])
        l0, p0 = lik_prior(L, M, tL, tM, i, poiA[i])
        l1, p1 = lik_prior(L2, M2, tL2, tM2, i, poiN)
        assert out["x"][i] == pytest.approx((l1 - l0) + (p1 - p0) + h, rel=1e-9, abs=1e-8), (i, kind[i])
        n_checked += 1
    assert n_checked > 60
    # round trip: remove the shift an add just inserted
    adds = [i for i in range(n) if kind[i] == 2 and out["ok"][i]]
    st2 = [((out["rates"][i], states[i][1], out["times"][i], states[i][3]) if side[i] else
            (states[i][0], out["rates"][i], states[i][2], out["times"][i])) for i in adds]
    back = evaluate_proposals(ds, st2, side[adds], 3, idx[adds] + 1, 0.5, 0.5, gamma_rate=g[adds], poi_lambda=lam[adds], poiA=poiA[adds])
    for k, i in enumerate(adds):
        r0 = states[i][0] if side[i] else states[i][1]
        np.testing.assert_allclose(back["rates"][k], r0, rtol=1e-11)
        assert back["hasting"][k] == pytest.approx(-out["hasting"][i], rel=1e-9, abs=1e-9)


def test_rate_multiplier_proposal_matches_the_oracle(device, metal_path):
    """update_multiplier_freq (:165-176) with an explicit mask and uniforms: proposed rates, Hastings term and the acceptance
    ratio the loop forms from per-lane differences against the oracle's absolute form, also for a heated chain (beta < 1)."""
    from literate_b200.engine import evaluate_proposals, LR_KMAX
    lin = O.read_lineages(metal_path)
    st, ds = _dataset(device, lin, 0)
    rng = np.random.default_rng(77)
    span = lin.end_time - lin.start_time
    states = random_states(rng, 64, lin.start_time, lin.end_time, kmax=9, rate_scale=0.3)
    n = len(states)
    side = rng.integers(0, 2, n)
    g = rng.gamma(2.0, 1.0, (n, 2)) + 0.05
    beta = np.where(rng.random(n) < 0.5, 1.0, rng.uniform(0.3, 1.0, n))
    on = (rng.random((n, LR_KMAX)) < 0.75).astype(np.int32)
    u = rng.random((n, LR_KMAX))
    out = evaluate_proposals(ds, states, side, 0, 0, 0.5, 0.5, gamma_rate=g, poi_lambda=1.0, beta=beta, poiA=0.0, mult_on=on, mult_u=u)
    for i, (L, M, tL, tM) in enumerate(states):
        r = L if side[i] else M
        rn, h = O.rate_multiplier_given(r, on[i, :len(r)], u[i, :len(r)])
        np.testing.assert_allclose(out["rates"][i], rn, rtol=1e-13)
        assert out["hasting"][i] == pytest.approx(h, rel=1e-11, abs=1e-13)
        L2, M2 = (rn, M) if side[i] else (L, rn)
        dl = O.loglik_state(L2, M2, tL, tM, st, 0) - O.loglik_state(L, M, tL, tM, st, 0)
        dp = O.state_prior(L2, M2, g[i], span, 0.0) - O.state_prior(L, M, g[i], span, 0.0)
        assert out["x"][i] == pytest.approx(beta[i] * dl + dp + h, rel=1e-9, abs=1e-8)
