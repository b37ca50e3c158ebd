"""GPU: the host-buffer entry points of the whole path -- engine.run_rjmcmc (one batch) and engine.Pipeline (a stream of
batches, copies of batch k+1 overlapped with the chains of batch k) -- give the same records as the step-by-step API."""
import numpy as np
import pytest

from oracle import literate_oracle as O
from literate_b200 import engine as E, synth

pytestmark = pytest.mark.gpu


def _batch(seed, n_rep=3, n=20_001):
    ts = np.empty((n_rep, n)); te = np.empty((n_rep, n))
    for r in range(n_rep):
        ts[r], te[r] = synth.syn_int(n, replicate=1000 * seed + r)
    return ts, te


def test_run_rjmcmc_equals_step_by_step(device):
    ts, te = _batch(1)
    rep = np.array([0, 1, 2, 0, 1, 2], dtype=np.int32)
    rec, stats = E.run_rjmcmc(device, ts, te, 6, 4001, 200, seed=9, first_bin=1800, n_bins=200, start_time=1800.0, end_time=2000.5, rep_of_chain=rep)
    st = device.bin_stats(ts, te, first_bin=1800, n_bins=200, end_time=2000.5)
    ds = E.Dataset(device, st, 0, 1800.0, 2000.5)
    want = E.Chains(ds, 6, 9, rep_of_chain=rep).run(4001, 200)
    assert np.array_equal(rec, want) and np.array_equal(stats.sp, st.sp) and np.array_equal(stats.br, st.br)
    for r in range(3):
        o = O.bin_stats_fast(ts[r], te[r])
        assert (st.sp[r] == o.sp).all() and (st.ex[r] == o.ex).all() and (st.br[r] == o.br).all()
    # each chain's likelihood is that of ITS replicate
    for c in range(6):
        L, M, tL, tM = E.record_to_state(rec[-1, c], 2000.5)
        o = O.bin_stats_fast(ts[rep[c]], te[rep[c]])
        assert rec[-1, c, E.REC_LIK] == pytest.approx(O.loglik_state(L, M, tL, tM, o, 0), rel=1e-10)


def test_pipeline_returns_every_batch_in_order():
    pipe = E.Pipeline(0)
    batches = [_batch(s) for s in (2, 3, 4, 5)]
    got = []
    for k, (ts, te) in enumerate(batches):
        prev = pipe.push(ts, te, 3, 3001, 250, seed=20 + k, first_bin=1800, n_bins=200, start_time=1800.0, end_time=2000.5,
                         rep_of_chain=np.arange(3, dtype=np.int32))
        assert (prev is None) == (k == 0)
        if prev is not None:
            got.append(prev)
    got.append(pipe.flush())
    assert pipe.flush() is None
    dev = pipe.dev_run
    for k, ((ts, te), (rec, stats)) in enumerate(zip(batches, got)):
        want, wst = E.run_rjmcmc(dev, ts, te, 3, 3001, 250, seed=20 + k, first_bin=1800, n_bins=200, start_time=1800.0, end_time=2000.5,
                                 rep_of_chain=np.arange(3, dtype=np.int32))
        assert np.array_equal(rec, want) and np.array_equal(stats.sp, wst.sp)
    pipe.close()
