"""GPU: K3 (the RJMCMC chains) through the C ABI.

State-level parity is deterministic (same state in -> same likelihood / prior out, checked in
test_gpu_likelihood.py and again here on the states the chains actually visit).  Chain-level parity is
distributional: the reference draws from NumPy's MT19937, the kernel from Philox, so posterior summaries
are compared with the committed summaries of 8 unmodified reference chains (tests/golden/posterior/*.json,
made by oracle/make_golden.py) using the between-chain spread as the Monte-Carlo error.
"""
import json
import os

import numpy as np
import pytest

from conftest import GOLD, golden_input
from oracle import literate_oracle as O
from literate_b200 import engine as E, synth

pytestmark = pytest.mark.gpu


def _setup(device, path, model=0, tbp=False, n_chains=64, seed=1, **cfg):
    lin = O.read_lineages(path, TBP=tbp)
    st = device.bin_stats(lin.ts, lin.te, only_dead=(model == 3), end_time=lin.end_time)
    ds = E.Dataset(device, st, model, lin.start_time, lin.end_time)
    ch = E.Chains(ds, n_chains, seed, E.default_config(model, **cfg))
    ost = O.BinStats(st.first_bin, st.sp[0], st.ex[0], st.br[0], None if st.ex_dead is None else st.ex_dead[0],
                     None if st.br_dead is None else st.br_dead[0])
    return lin, ost, ds, ch


def test_initial_state_matches_reference_bookkeeping(device):
    lin, st, ds, ch = _setup(device, golden_input("example_dataTAD.txt"), n_chains=8)
    rec = ch.state()
    for r in rec:
        L, M, tL, tM = E.record_to_state(r, lin.end_time)
        assert len(L) == 1 and len(M) == 1 and r[E.REC_IT] == 0
        assert r[E.REC_LIK] == pytest.approx(O.loglik_state(L, M, tL, tM, st, 0), rel=1e-10)
        # initial prior uses prior_gamma's default rate 2 (LiteRateForward.py:227), not Gamma_rate = 1
        want = O.rates_prior(L) + O.rates_prior(M) + 2 * O.poisson_prior(1, 1)
        assert r[E.REC_PRIOR] == pytest.approx(want, rel=1e-10)
        assert r[E.REC_GL] == 1.0 and r[E.REC_GM] == 1.0 and r[E.REC_POI] == 1.0 and r[E.REC_POI_INIT] == 1.0
    # Gamma(2, scale 2) initial rates: mean 4, var 8
    lin, st, ds, big = _setup(device, golden_input("example_dataTAD.txt"), n_chains=4096, seed=5)
    r0 = big.state()
    for col in (E.REC_L, E.REC_M):
        x = r0[:, col]
        assert abs(x.mean() - 4.0) < 5 * np.sqrt(8 / 4096) and abs(x.var() - 8.0) < 1.5


@pytest.mark.parametrize("model", [0, 1, 2, 3])
def test_visited_states_are_consistent(device, model, metal_path):
    """Every logged state: likelihood and prior recomputed by the oracle from the logged rates/times agree,
    shift spacing obeys the guard (:290), K matches, adequacy matches calculate_r_squared."""
    lin, st, ds, ch = _setup(device, metal_path if model != 1 else golden_input("example_dataTAD.txt"), model=model, n_chains=16, seed=3)
    recs = ch.run(20000, 500)
    assert recs.shape[0] == 40
    emp_b, emp_d = st.sp / st.br, st.ex / st.br
    seen_k = set()
    n_stale = n_seen = 0
    for s in range(recs.shape[0]):
        for c in range(0, 16, 5):
            r = recs[s, c]
            L, M, tL, tM = E.record_to_state(r, lin.end_time)
            seen_k.add(len(L))
            assert r[E.REC_IT] == s * 500
            assert np.min(np.diff(tL)) > 1 and np.min(np.diff(tM)) > 1
            assert r[E.REC_LIK] == pytest.approx(O.loglik_state(L, M, tL, tM, st, model), rel=1e-10)
            assert r[E.REC_LAVG] == pytest.approx(L.mean(), rel=1e-12) and r[E.REC_MAVG] == pytest.approx(M.mean(), rel=1e-12)
            if s > 4:   # after the transient of :227 the stored prior is the prior of the stored state:
                # rates + shift-time prior under the logged hyper-parameters + the stored Poisson term priorPoiA (record
                # slot 15), which is the Poisson prior of SOME earlier (K_l, K_m, poisson rate): stale on purpose (:300-304, :319)
                want = O.state_prior(L, M, [r[E.REC_GL], r[E.REC_GM]], lin.end_time - lin.start_time, r[E.REC_POIA])
                assert r[E.REC_PRIOR] == pytest.approx(want, rel=1e-10, abs=1e-10)
                fresh = O.poisson_prior(len(L), r[E.REC_POI]) + O.poisson_prior(len(M), r[E.REC_POI])
                n_stale += int(abs(r[E.REC_POIA] - fresh) > 1e-9 * max(1.0, abs(fresh)))
                n_seen += 1
            iL = O.rate_index(np.floor(tL) if len(tL) > 2 else tL, st.n_bins)
            iM = O.rate_index(np.floor(tM) if len(tM) > 2 else tM, st.n_bins)
            np.testing.assert_allclose(r[E.REC_ADQ:E.REC_ADQ + 3], O.adequacy_closed_form(emp_b, emp_d, L[iL], M[iM]), rtol=1e-10)
            np.testing.assert_allclose(r[E.REC_ADQ:E.REC_ADQ + 3], O.adequacy(emp_b, emp_d, L[iL], M[iM]), rtol=1e-8)    # np.linalg.lstsq's own rounding
    assert len(seen_k) > 1      # the sampler does jump between dimensions
    assert 0 < n_stale < n_seen  # quirk (ii) is live: the stored Poisson term is stale in some records and fresh in others
    cnt = ch.counters()
    assert (cnt[:, 0] == 20000).all() and (cnt[:, 7] == 0).all()
    frac = cnt.sum(0) / cnt[:, 0].sum()
    assert abs(frac[5] - 0.199) < 0.01 and abs(frac[6] - 0.001) < 0.0005     # RJ 19.9 %, Gibbs 0.1 % (:274, :281)


def test_determinism_and_sharding_invariance(device):
    """Same (seed, chain id) -> same chain, whatever the launch split or the population it is part of."""
    path = golden_input("example_dataTAD.txt")
    lin, st, ds, a = _setup(device, path, n_chains=32, seed=11)
    ra = a.run(3000, 100)
    b = E.Chains(ds, 32, 11)
    rb = np.concatenate([b.run(1000, 100), b.run(1500, 100), b.run(500, 100)])
    assert np.array_equal(ra, rb)
    c = E.Chains(ds, 8, 11, chain_id0=16)          # chains 16..23 as a shard of their own
    rc = c.run(3000, 100)
    assert np.array_equal(ra[:, 16:24], rc)
    d = E.Chains(ds, 32, 12)
    assert not np.array_equal(ra, d.run(3000, 100))


def test_loop_builds_give_identical_chains(device, metal_path):
    """The latency-optimised and the compact build of the chain loop (lr_chain_config.loop_variant) are the same
    arithmetic: identical records, bit for bit."""
    lin, st, ds, a = _setup(device, metal_path, n_chains=22, seed=5, loop_variant=1)
    b = E.Chains(ds, 22, 5, E.default_config(0, loop_variant=2))
    c = E.Chains(ds, 22, 5, E.default_config(0, loop_variant=3))       # four chains per CTA: 22 leaves a ragged last CTA
    ra, rb, rc = a.run(30000, 250), b.run(30000, 250), c.run(30000, 250)
    assert np.array_equal(ra, rb) and np.array_equal(ra, rc)
    assert np.array_equal(a.counters(), b.counters()) and np.array_equal(a.counters(), c.counters())


@pytest.mark.parametrize("team_w,lead,nobail", [(4, 1, 1), (8, 2, 1), (16, 7, 1), (10, 3, 1), (8, 5, 0), (4, 4, 0), (16, 2, 0), (12, 4, 0)])
def test_speculative_team_build_gives_the_same_chains(device, metal_path, team_w, lead, nobail, monkeypatch):
    """loop_variant 4: W warps evaluate consecutive iterations of ONE chain ahead of time against the current state and only
    the first state-changing one commits (csrc/k3_team.cuh).  Same draws, same arithmetic, same state: the records, final
    states and event counters must be those of the compact build, bit for bit -- on a table where a third of the iterations
    change the state (metal bands, many rollbacks) and on one where almost none does, split over ragged launches."""
    monkeypatch.setenv("LR_TEAM_W", str(team_w))
    monkeypatch.setenv("LR_TEAM_LEAD", str(lead))
    # nobail = 1: the team runs every iteration however often the state changes; 0 (the default): a team whose chain changes
    # state more than once in 8 iterations stops and the continuation pass (one-chain CTAs) finishes the launch
    monkeypatch.setenv("LR_TEAM_NOBAIL", str(nobail))
    lin, st, ds, a = _setup(device, metal_path, n_chains=11, seed=5, loop_variant=2)
    b = E.Chains(ds, 11, 5, E.default_config(0, loop_variant=4))
    ra, rb = a.run(30000, 250), b.run(30000, 250)
    assert np.array_equal(ra, rb)
    assert np.array_equal(a.counters(), b.counters())
    assert np.array_equal(a.state(), b.state())
    for k, n in enumerate([1, 7, 8, 9, 23, 64, 65, 1000, 3, 1025, 4096, 5]):
        s = [0, 1, 5][k % 3]
        ra, rb = a.run(n, s), b.run(n, s)
        if s:
            assert np.array_equal(ra, rb), (k, n, s)
        assert np.array_equal(a.state(), b.state()), (k, n)
    assert np.array_equal(a.counters(), b.counters())
    # every sampler configuration and model once, from the initial state (quirk i: the first iterations run in the absolute form)
    for model, kw in [(0, dict(const_death_rate=1)), (0, dict(const_rates=1)), (0, dict(Poisson_prior=2.0, use_rate_HP=0)),
                      (1, {}), (2, {}), (3, {}), (0, dict(real_move_shift=1)), (0, dict(beta=0.3))]:
        path = golden_input("example_dataTAD.txt")
        lin, st, ds, a = _setup(device, path, model=model, n_chains=5, seed=77, loop_variant=2, **kw)
        b = E.Chains(ds, 5, 77, E.default_config(model, loop_variant=4, **kw))
        ra, rb = a.run(6000, 100), b.run(6000, 100)
        assert np.array_equal(ra, rb), (model, kw)
        assert np.array_equal(a.counters(), b.counters()), (model, kw)
    if not nobail:
        # on the 75-lineage table 40 % of the iterations change the state: every team must have handed its chain over
        lin, st, ds, b = _setup(device, golden_input("example_dataTAD.txt"), n_chains=5, seed=78, loop_variant=4)
        b.run(50000)
        t = b.team_stats()
        assert (b.counters()[:, 0] == 50000).all() and (t[:, 0] < 0.2 * 50000).all() and (t[:, 0] > 0).all()


def test_loop_builds_identical_for_ragged_launch_lengths(device):
    """The producer/consumer ring of the specialised build hands iterations over in batches of 8: launches of 1, 7, 8, 9, ...
    iterations (partial first/last batches, sampling on and off) must still be the compact build's chain, bit for bit."""
    lin, st, ds, a = _setup(device, golden_input("example_dataTAD.txt"), n_chains=9, seed=21, loop_variant=1)
    b = E.Chains(ds, 9, 21, E.default_config(0, loop_variant=2))
    c = E.Chains(ds, 9, 21, E.default_config(0, loop_variant=3))
    for k, n in enumerate([1, 7, 8, 9, 23, 64, 65, 1000, 3, 1025, 4096, 5]):
        s = [0, 1, 5][k % 3]
        ra, rb, rc = a.run(n, s), b.run(n, s), c.run(n, s)
        if s:
            assert np.array_equal(ra, rb) and np.array_equal(ra, rc), (k, n, s)
        assert np.array_equal(a.state(), b.state()) and np.array_equal(a.state(), c.state()), (k, n)
    assert np.array_equal(a.counters(), b.counters()) and np.array_equal(a.counters(), c.counters())


def test_many_rates_per_side_and_capacity(device, metal_path):
    """States with up to LR_KMAX = 30 rates per side: likelihood parity holds there, add-shift proposals at capacity are
    rejected and counted, and the chain keeps running."""
    lin, st, ds, ch = _setup(device, metal_path, n_chains=6, seed=8)
    rng = np.random.default_rng(12)
    rec = ch.state()
    K = 30
    for r in rec:
        for off_r, off_t, kcol in ((E.REC_L, E.REC_TL, E.REC_KL), (E.REC_M, E.REC_TM, E.REC_KM)):
            r[kcol] = K
            r[off_r:off_r + K] = rng.gamma(2.0, 0.2, K)
            r[off_t + 1:off_t + K] = lin.start_time + 1.05 * np.arange(1, K) + rng.uniform(0, 0.04, K - 1)   # spacing > 1
    ch.set_state(rec)
    got = ch.state()
    for r in got:
        L, M, tL, tM = E.record_to_state(r, lin.end_time)
        assert len(L) == K and len(M) == K
        assert r[E.REC_LIK] == pytest.approx(O.loglik_state(L, M, tL, tM, st, 0), rel=1e-10)
    recs = ch.run(4000, 100)
    cnt = ch.counters()
    assert cnt[:, 7].sum() > 0                     # add-shift at capacity: rejected and counted
    assert (recs[:, :, E.REC_KL] <= K).all() and (recs[:, :, E.REC_KM] <= K).all() and np.isfinite(recs[:, :, E.REC_LIK]).all()
    r = recs[-1, 0]
    L, M, tL, tM = E.record_to_state(r, lin.end_time)
    assert r[E.REC_LIK] == pytest.approx(O.loglik_state(L, M, tL, tM, st, 0), rel=1e-10)


def test_four_thousand_chains_shard_invariance(device, metal_path):
    """BASELINE cfg4 population size (4096 chains, compact build): any shard of the population reproduces its slice bit for
    bit, whatever build the shard size selects, and tempered swap rounds keyed by global ladder ids agree as well."""
    lin, st, ds, whole = _setup(device, metal_path, n_chains=4096, seed=404)
    parts = [(0, 256), (256, 1024), (1280, 2816)]              # one chain per CTA, four chains per CTA, compact build
    shards = [E.Chains(ds, n, 404, chain_id0=c0) for c0, n in parts]
    from literate_b200 import parallel as P
    beta = np.tile(P.temperature_ladder(8, 0.1), 512)
    whole.set_beta(beta)
    for (c0, n), s in zip(parts, shards):
        s.set_beta(beta[c0:c0 + n])
    rw = whole.run(1501, 500)
    for (c0, n), s in zip(parts, shards):
        assert np.array_equal(s.run(1501, 500), rw[:, c0:c0 + n])
    for rnd in range(3):
        whole.swap_step(8, rnd); whole.run(300)
        for s in shards:
            s.swap_step(8, rnd); s.run(300)
    sw = whole.state()
    for (c0, n), s in zip(parts, shards):
        assert np.array_equal(s.state(), sw[c0:c0 + n])
    assert whole.counters()[:, 9].sum() > 0


def test_checkpoint_roundtrip(device):
    path = golden_input("example_dataTAD.txt")
    lin, st, ds, a = _setup(device, path, n_chains=4, seed=2)
    a.run(5000)
    snap = a.state()
    b = E.Chains(ds, 4, 2)
    b.set_state(snap)
    got = b.state()
    np.testing.assert_array_equal(got[:, E.REC_L:], snap[:, E.REC_L:])
    np.testing.assert_allclose(got[:, E.REC_LIK], snap[:, E.REC_LIK], rtol=1e-13)


Z_MAX = 3.5          # standard errors of the difference of the two chain-population means (fixed seeds: the tests are deterministic)


def _summaries(recs, lin, burnin=0.2):
    ns = recs.shape[0]
    b = int(burnin * ns)
    out = []
    for c in range(recs.shape[1]):
        r = recs[b:, c]
        rowsL = [np.concatenate([x[E.REC_L:E.REC_L + int(x[E.REC_KL])], x[E.REC_TL + 1:E.REC_TL + int(x[E.REC_KL])]]) for x in r]
        rowsM = [np.concatenate([x[E.REC_M:E.REC_M + int(x[E.REC_KM])], x[E.REC_TM + 1:E.REC_TM + int(x[E.REC_KM])]]) for x in r]
        out.append({"K_l": r[:, E.REC_KL].mean(), "K_m": r[:, E.REC_KM].mean(), "lik": r[:, E.REC_LIK].mean(), "lik_var": r[:, E.REC_LIK].var(),
                    "K_l_pmf": np.bincount(r[:, E.REC_KL].astype(int), minlength=E.LR_KMAX + 2) / len(r),
                    "K_m_pmf": np.bincount(r[:, E.REC_KM].astype(int), minlength=E.LR_KMAX + 2) / len(r),
                    "lam": r[:, E.REC_LAVG].mean(), "mu": r[:, E.REC_MAVG].mean(),
                    "gL": r[:, E.REC_GL].mean(), "gM": r[:, E.REC_GM].mean(), "poi": r[:, E.REC_POI].mean(),
                    "birth": O.marginal_rates(rowsL, lin.end_time, lin.start_time, 0).mean(0),
                    "death": O.marginal_rates(rowsM, lin.end_time, lin.start_time, 0).mean(0)})
    return out


def _pmf_rows(ref, key):
    rows = np.zeros((len(ref), E.LR_KMAX + 2))
    for i, r in enumerate(ref):
        tot = sum(r[key].values())
        for k, v in r[key].items():
            rows[i, int(k)] = v / tot
    return rows


def _compare_with_reference(mine, ref, what=("K_l", "K_m", "lik", "lambda_avg", "mu_avg", "birth", "death")):
    """Two populations of chains of the same length from the same initial distribution.  Per chain: means over the
    post-burn-in samples; per population: mean and standard error over chains.  Every mean, every per-bin marginal rate, the
    variance of the log-likelihood trace and every cell of the K_l / K_m distributions must agree within Z_MAX standard errors;
    the K distributions also as a whole (chi-square over the cells with at least 2 % of the mass)."""
    from scipy import stats as sps

    def kmean(pmf):
        tot = sum(pmf.values())
        return sum(int(k) * v for k, v in pmf.items()) / tot

    def compare(name, a, b, floor=0.0):
        a, b = np.asarray(a, float), np.asarray(b, float)
        se = np.sqrt(a.var(0, ddof=1) / len(a) + b.var(0, ddof=1) / len(b)) + floor
        if np.all(se == 0):                      # a constant on both sides (hyper-parameters that are never resampled)
            assert np.allclose(a.mean(0), b.mean(0), rtol=1e-12), (name, a.mean(0), b.mean(0))
            return
        z = np.abs(a.mean(0) - b.mean(0)) / se
        assert np.all(z < Z_MAX), (name, float(np.max(z)), a.mean(0), b.mean(0))

    def compare_pmf(name, a, b):
        pooled = 0.5 * (a.mean(0) + b.mean(0))
        cells = np.nonzero(pooled >= 0.02)[0]
        if len(cells) < 2:
            assert np.allclose(a.mean(0)[cells], b.mean(0)[cells], atol=0.02), (name, a.mean(0), b.mean(0))
            return
        se = np.sqrt(a[:, cells].var(0, ddof=1) / len(a) + b[:, cells].var(0, ddof=1) / len(b)) + 1e-9
        z = (a[:, cells].mean(0) - b[:, cells].mean(0)) / se
        assert np.all(np.abs(z) < Z_MAX), (name, cells, z)
        # cells sum to (nearly) one: len(cells) - 1 degrees of freedom, at the 0.1 % level
        assert float(np.sum(z ** 2)) < sps.chi2.ppf(0.999, len(cells) - 1) * len(cells) / (len(cells) - 1), (name, cells, z)

    if "K_l" in what:
        compare("K_l", [m["K_l"] for m in mine], [kmean(r["K_l"]) for r in ref], floor=1e-9)
        compare_pmf("K_l pmf", np.array([m["K_l_pmf"] for m in mine]), _pmf_rows(ref, "K_l"))
    if "K_m" in what:
        compare("K_m", [m["K_m"] for m in mine], [kmean(r["K_m"]) for r in ref], floor=1e-9)
        compare_pmf("K_m pmf", np.array([m["K_m_pmf"] for m in mine]), _pmf_rows(ref, "K_m"))
    compare("lik", [m["lik"] for m in mine], [r["lik_mean"] for r in ref])
    compare("lik_var", [m["lik_var"] for m in mine], [r["lik_var"] for r in ref])
    compare("lambda_avg", [m["lam"] for m in mine], [r["lambda_avg"] for r in ref])
    compare("mu_avg", [m["mu"] for m in mine], [r["mu_avg"] for r in ref])
    compare("birth", [m["birth"] for m in mine], [r["birth_rate_mean"] for r in ref], floor=1e-6)
    compare("death", [m["death"] for m in mine], [r["death_rate_mean"] for r in ref], floor=1e-6)
    # hyper-parameters drawn by the Gibbs steps (:99-108, :210-213): integer-shape Gamma and Marsaglia-Tsang samplers
    compare("gamma_rate_hp_BI", [m["gL"] for m in mine], [r["gamma_hp_l"] for r in ref])
    compare("gamma_rate_hp_D", [m["gM"] for m in mine], [r["gamma_hp_m"] for r in ref])
    compare("poisson_rate_hp", [m["poi"] for m in mine], [r["poisson_hp"] for r in ref])


@pytest.mark.parametrize("data,n_iter,s", [("example_tad", 300000, 100), ("metal_bands", 400000, 200)])
def test_posterior_matches_reference_chains(device, data, n_iter, s, metal_path):
    """128 GPU chains vs 32 unmodified reference chains (same length, same sampling, 20 % burn-in): the means of
    K_l, K_m, log-likelihood (and its variance), mean rates, every per-bin marginal rate, the hyper-parameters and the
    K_l / K_m distributions agree within 3.5 standard errors of the difference of the two chain-population means."""
    with open(os.path.join(GOLD, "posterior", data + ".json")) as fh:
        ref = json.load(fh)["chains"]
    assert len(ref) >= 32
    path = golden_input("example_dataTAD.txt") if data == "example_tad" else metal_path
    lin, st, ds, ch = _setup(device, path, n_chains=128, seed=2026)
    recs = ch.run(n_iter + 1, s)
    _compare_with_reference(_summaries(recs, lin), ref)


def test_posterior_on_the_bench_statistics_matches_the_oracle_chains(device):
    """The workload the metric is quoted on: 1M synthetic lineages x 200 one-year bins (K_l ~ 11, K_m = 1).  256 device chains
    (the speculative team build, as in bench.py) against 32 chains of the byte-pinned oracle on the same statistics
    (tests/golden/posterior/syn_int_200.json, `oracle/make_golden.py posterior_syn`), 400 001 iterations each, 50 % burn-in:
    both populations start from the reference's initial distribution and have the same finite length, so their per-chain
    summaries are draws from one distribution even where single chains have not converged."""
    with open(os.path.join(GOLD, "posterior", "syn_int_200.json")) as fh:
        fx = json.load(fh)
    ref = fx["chains"]
    ts, te = synth.syn_int(1_000_000, 0)
    st = device.bin_stats(ts, te, death_jitter=0.5)
    want = O.bin_stats_fast(ts, te)
    assert np.array_equal(st.sp[0], want.sp) and np.array_equal(st.ex[0], want.ex) and np.array_equal(st.br[0], want.br)
    lin = O.Lineages(ts=ts, te=te, start_time=float(ts.min()), end_time=float(te.max()), true_root_age=float(ts.min()))
    ds = E.Dataset(device, st, 0, lin.start_time, lin.end_time)
    ch = E.Chains(ds, 256, seed=4242)
    recs = ch.run(ref[0]["n_iterations"], ref[0]["s_freq"])
    assert (ch.team_stats()[:, 4] > 0.9 * ref[0]["n_iterations"]).all()         # the teams did the work
    _compare_with_reference(_summaries(recs, lin, burnin=fx["burnin"]), ref)


FLAG_SETS = {   # tests/golden/posterior/<tag>.json, made by `oracle/make_golden.py posterior_flags` from the unmodified reference
    "tad_constdeath": dict(model=0, const_death_rate=1),           # -const_death_rate 1 (:243-247)
    "tad_constrates": dict(model=0, const_rates=1),                # -const_rates 1 (:274, :281)
    "tad_fixedpoi_nohp": dict(model=0, Poisson_prior=2.0, use_rate_HP=0),    # -Poisson_prior 2 -use_rate_HP 0 (:220-221, :283-286)
    "tad_keiding": dict(model=2),                                  # -model_BDI 2 (:137-148)
    "tad_immigration": dict(model=1),                              # -model_BDI 1 (:151-153)
    "tad_keiding_dead": dict(model=3),                             # -model_BDI 3 (:529-549)
}


@pytest.mark.parametrize("tag", sorted(FLAG_SETS))
def test_posterior_matches_reference_for_every_sampler_configuration(device, tag):
    """The same comparison for each non-default flag set: constant death rate / constant rates (RJ disabled on one or both
    sides), fixed Poisson prior without rate hyper-prior (no Gibbs moves), and the three other likelihoods."""
    p = os.path.join(GOLD, "posterior", tag + ".json")
    if not os.path.exists(p):
        pytest.skip("fixture not generated")
    with open(p) as fh:
        ref = json.load(fh)["chains"]
    cfg = dict(FLAG_SETS[tag])
    model = cfg.pop("model")
    lin, st, ds, ch = _setup(device, golden_input("example_dataTAD.txt"), model=model, n_chains=128, seed=77, **cfg)
    recs = ch.run(ref[0]["n_iterations"], ref[0]["s_freq"])
    mine = _summaries(recs, lin)
    if tag == "tad_constrates":
        assert all(m["K_l"] == 1 and m["K_m"] == 1 for m in mine)
    if tag == "tad_constdeath":
        assert all(m["K_m"] == 1 for m in mine)
    _compare_with_reference(mine, ref)
    cnt = ch.counters().sum(0)
    if tag == "tad_constrates":
        assert cnt[5] == 0 and abs(cnt[6] / cnt[0] - 0.2) < 0.005          # every r0 >= .8 goes to the Gibbs branch (:281)


def test_prior_only_matches_the_oracle_chain(device):
    """All sufficient statistics zero: the likelihood vanishes and the chains sample the prior that the proposals, the
    Hastings/Jacobian terms and the hyper-prior Gibbs steps define.  64 GPU chains against 8 oracle chains (the restatement
    that reproduces the reference byte for byte; tests/golden/posterior/prior_only.json, `make_golden.py prior_only`):
    number-of-rates distribution and the Poisson hyper-parameter agree within Monte-Carlo error."""
    with open(os.path.join(GOLD, "posterior", "prior_only.json")) as fh:
        ref = json.load(fh)["chains"]
    lin = O.read_lineages(golden_input("example_dataTAD.txt"))
    nb = int(lin.end_time) - int(lin.start_time)
    z = np.zeros((1, nb))
    ds = E.Dataset(device, E.BinStats(int(lin.start_time), z.astype(np.int64), z.astype(np.int64), z), 0, lin.start_time, lin.end_time)
    ch = E.Chains(ds, 64, seed=606)
    recs = ch.run(ref[0]["n_iterations"], ref[0]["s_freq"])
    assert np.all(recs[:, :, E.REC_LIK] == 0.0)
    post = recs[recs.shape[0] // 5:]

    def frac(pmf, k):
        return pmf.get(str(k), 0) / sum(pmf.values())

    def close(name, a, b):
        a, b = np.asarray(a, float), np.asarray(b, float)
        se = np.sqrt(a.var(ddof=1) / len(a) + b.var(ddof=1) / len(b))
        assert abs(a.mean() - b.mean()) < 4.5 * se, (name, a.mean(), b.mean(), se)

    for col, key in ((E.REC_KL, "K_l"), (E.REC_KM, "K_m")):
        close(key + " mean", post[:, :, col].mean(0), [sum(int(k) * v for k, v in r[key].items()) / sum(r[key].values()) for r in ref])
        for k in (1, 2, 3):
            close("%s = %d" % (key, k), (post[:, :, col] == k).mean(0), [frac(r[key], k) for r in ref])
    close("poisson_hp", post[:, :, E.REC_POI].mean(0), [r["poisson_hp"] for r in ref])


def test_gibbs_poisson_rate_draws_are_gamma(device):
    """With -const_rates 1 the dimension never changes (K_l = K_m = 1) and a fifth of the iterations redraw the Poisson rate
    from Gamma(2 + K_l + K_m, scale 1/3) (get_post_rj_HP, :99-108): the logged values must be i.i.d. draws of Gamma(4, 1/3)
    (Kolmogorov-Smirnov against scipy), which exercises the device's integer-shape Gamma sampler and the Philox streams."""
    import scipy.stats
    lin, st, ds, ch = _setup(device, golden_input("example_dataTAD.txt"), n_chains=32, seed=99, const_rates=1)
    recs = ch.run(20001, 100)
    x = recs[5:, :, E.REC_POI].ravel()                 # 50 samples apart: a fresh draw almost surely (1 - 0.8^100)
    assert len(np.unique(x)) > 0.99 * len(x)
    ks = scipy.stats.kstest(x, scipy.stats.gamma(4.0, scale=1.0 / 3.0).cdf)
    assert ks.pvalue > 1e-3, ks
    assert abs(x.mean() - 4.0 / 3.0) < 5 * np.sqrt((4.0 / 9.0) / len(x))
    # independent streams: chains are uncorrelated
    c = np.corrcoef(recs[5:, :8, E.REC_POI].T)
    assert np.max(np.abs(c - np.eye(8))) < 0.25


def test_team_build_under_load_many_chains_and_ragged_launches(device, monkeypatch):
    """Two teams on every SM (296 chains, the occupancy the bench runs at) through 16 launches of random lengths and sampling
    periods: records, states, counters of the compact build, bit for bit -- the protocol's races would show here first."""
    monkeypatch.setenv("LR_TEAM_NOBAIL", "1")
    ts, te = synth.syn_int(200_000, 3)
    st = device.bin_stats(ts, te)
    ds = E.Dataset(device, st, 0, float(ts.min()), float(te.max()))
    a = E.Chains(ds, 296, 9, E.default_config(0, loop_variant=2))
    b = E.Chains(ds, 296, 9, E.default_config(0, loop_variant=4))
    rng = np.random.default_rng(4)
    for k in range(16):
        n = int(rng.integers(1, 6000)); s = int(rng.choice([0, 1, 3, 64, 1000]))
        ra, rb = a.run(n, s), b.run(n, s)
        if s:
            assert np.array_equal(ra, rb), (k, n, s)
    assert np.array_equal(a.state(), b.state()) and np.array_equal(a.counters(), b.counters())
    t = b.team_stats().sum(0)
    assert t[0] > 0 and t[1] > 0          # commits and rollbacks did happen
