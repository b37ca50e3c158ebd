"""GPU: K4, the Keiding log-likelihood evaluated directly over the lineages (validation path, lr_loglik_direct), against
(1) the production path K1 + K2 on the same states and (2) the oracle.  Tolerance 1e-10 relative (north star)."""
import numpy as np
import pytest
import torch

from conftest import golden_input, random_states
from oracle import literate_oracle as O
from literate_b200 import engine as E, synth

pytestmark = pytest.mark.gpu


def _per_bin(states, n_bins):
    lam = np.empty((len(states), n_bins)); mu = np.empty((len(states), n_bins))
    for i, (L, M, tL, tM) in enumerate(states):
        lam[i] = L[O.rate_index(np.floor(tL) if len(tL) > 2 else tL, n_bins)]
        mu[i] = M[O.rate_index(np.floor(tM) if len(tM) > 2 else tM, n_bins)]
    return lam, mu


def _check(device, ts, te, n_states, seed, oracle_loop=True):
    first, nb = E.window(ts, te)
    start, end = float(ts.min()), float(te.max())
    rng = np.random.default_rng(seed)
    states = random_states(rng, n_states, start, end, kmax=7, rate_scale=0.2)
    lam, mu = _per_bin(states, nb)
    d = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
    got = device.loglik_direct_device(d(ts), d(te), first, nb, d(lam), d(mu)).cpu().numpy()
    st = device.bin_stats(ts, te, death_jitter=0.0)
    ds = E.Dataset(device, st, 2, start, end)
    binned = ds.evaluate(states)["lik"]
    np.testing.assert_allclose(got, binned, rtol=1e-10)
    if oracle_loop:
        ost = O.BinStats(first, st.sp[0], st.ex[0], st.br[0])
        want = np.array([O.loglik_keiding(lam[i], mu[i], ost, False) for i in range(n_states)])
        np.testing.assert_allclose(got, want, rtol=1e-10)
    return got


def test_shipped_tables(device, metal_path):
    for path, tbp in ((golden_input("example_dataTAD.txt"), False), (golden_input("example_dataTBP.txt"), True), (metal_path, False)):
        lin = O.read_lineages(path, TBP=tbp)
        _check(device, lin.ts, lin.te, 21, 5)       # 21 states: two full groups of 8 and a ragged one


def test_degenerate_lineages(device):
    # zero time at risk, te < ts, NaN rows, lineages born before / dying after / outside the window
    ts = np.array([5.0, 7.0, 9.0, 6.5, 8.0, 5.0, np.nan, 2.0, 11.9, 30.0, 5.25]); te = np.array([12.5, 7.0, 6.0, 6.5, np.nan, 5.5, 9.5, 8.5, 40.0, 31.0, 5.75])
    first, nb = 5, 7
    rng = np.random.default_rng(2)
    lam = rng.gamma(2.0, 0.2, (3, nb)); mu = rng.gamma(2.0, 0.2, (3, nb))
    d = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
    got = device.loglik_direct_device(d(ts), d(te), first, nb, d(lam), d(mu)).cpu().numpy()
    with np.errstate(invalid="ignore"):
        rows = [O.events_in_bin(ts, te, first + j, first + j + 1) for j in range(nb)]
    sp, ex, br = (np.array(x, dtype=float) for x in zip(*rows))
    want = (np.log(lam) * sp - lam * br + np.log(mu) * ex - mu * br).sum(1)
    np.testing.assert_allclose(got, want, rtol=1e-12)


def test_one_million_lineages_integer_and_real(device):
    """Full size: the direct pass over 1M lineages and the binned path agree to 1e-10 on 64 states, for integer-year and
    real-valued tables; a second evaluation returns the same bits (fixed-order reduction)."""
    for real in (0, 1):
        ts, te = synth.syn_real(1_000_000) if real else synth.syn_int(1_000_000)
        a = _check(device, ts, te, 64, 9 + real, oracle_loop=False)
        b = _check(device, ts, te, 64, 9 + real, oracle_loop=False)
        assert np.array_equal(a, b)
