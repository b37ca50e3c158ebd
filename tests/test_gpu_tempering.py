"""GPU: tempered ensembles (Metropolis-coupled chains).  The reference has no tempering (SURVEY A-15), so the checks are
self-consistency ones: a swap round permutes the temperatures of a ladder, its outcome does not depend on how the
ensemble is sharded over devices, and the cold chains of a tempered ensemble sample the same posterior as plain chains
(whose parity with the reference is established in test_gpu_chains.py)."""
import numpy as np
import pytest
import torch

from conftest import golden_input
from oracle import literate_oracle as O
from literate_b200 import engine as E, parallel as P

pytestmark = pytest.mark.gpu


def _dataset(device):
    lin = O.read_lineages(golden_input("example_dataTAD.txt"))
    st = device.bin_stats(lin.ts, lin.te)
    return lin, E.Dataset(device, st, 0, lin.start_time, lin.end_time)


def test_swap_round_permutes_temperatures_and_counts(device):
    lin, ds = _dataset(device)
    T, n = 4, 64
    ch = E.Chains(ds, n, seed=7)
    beta = np.tile(P.temperature_ladder(T, 0.3), n // T)
    ch.set_beta(beta)
    ch.run(2000)
    prev = beta.copy()
    for rnd in range(6):
        ch.swap_step(T, rnd)
        now = ch.state()[:, E.REC_BETA]
        assert np.array_equal(np.sort(now.reshape(-1, T), axis=1), np.sort(beta.reshape(-1, T), axis=1))   # a permutation per ladder
        # only temperature neighbours of this round's parity exchanged
        for lad in range(n // T):
            a, b = prev[lad * T:(lad + 1) * T], now[lad * T:(lad + 1) * T]
            ra = np.argsort(np.argsort(-a)); rb = np.argsort(np.argsort(-b))
            moved = np.nonzero(ra != rb)[0]
            for i in moved:
                lo = min(ra[i], rb[i])
                assert abs(int(ra[i]) - int(rb[i])) == 1 and lo % 2 == rnd % 2
        prev = now
        ch.run(500)
    cnt = ch.counters()
    assert cnt[:, 8].sum() > 0 and 0 < cnt[:, 9].sum() <= cnt[:, 8].sum()
    assert cnt[:, 8].sum() % 2 == 0 and cnt[:, 9].sum() % 2 == 0          # both members of a pair count
    # states (rates, times) never move between chains: only beta does
    ch2 = E.Chains(ds, n, seed=7); ch2.set_beta(beta); ch2.run(2000)
    s_before = ch2.state()
    ch2.swap_step(T, 0)
    s_after = ch2.state()
    assert np.array_equal(s_before[:, E.REC_L:], s_after[:, E.REC_L:])


def test_swap_outcome_is_independent_of_sharding(device):
    """Two shards of 8 chains each, fed the gathered (lik, beta) table, take the decisions the 16-chain population takes
    (the multi-rank path of parallel.tempered_swap, emulated on one device; ladders of 4 and a ladder that spans the shards)."""
    lin, ds = _dataset(device)
    for T in (4, 16):
        n = 16
        beta = np.tile(P.temperature_ladder(T, 0.25), n // T)
        whole = E.Chains(ds, n, seed=3); whole.set_beta(beta); whole.run(1500)
        a = E.Chains(ds, 8, seed=3, chain_id0=0); a.set_beta(beta[:8]); a.run(1500)
        b = E.Chains(ds, 8, seed=3, chain_id0=8); b.set_beta(beta[8:]); b.run(1500)
        for rnd in range(4):
            whole.swap_step(T, rnd)
            ia = a.swap_info_device(torch.empty((8, 2), dtype=torch.float64, device="cuda"))
            ib = b.swap_info_device(torch.empty((8, 2), dtype=torch.float64, device="cuda"))
            device.sync()
            torch.cuda.synchronize()
            table = torch.cat([ia, ib])
            a.swap_apply_device(table, 0, T, rnd); b.swap_apply_device(table, 8, T, rnd)
            device.sync()
            got = np.concatenate([a.state()[:, E.REC_BETA], b.state()[:, E.REC_BETA]])
            assert np.array_equal(got, whole.state()[:, E.REC_BETA]), (T, rnd)
            whole.run(300); a.run(300); b.run(300)
        sw, sa, sb = whole.state(), a.state(), b.state()
        assert np.array_equal(sw, np.concatenate([sa, sb]))
        if T == 4:
            # shards that hold whole ladders may use the local swap_step: ladders are keyed by GLOBAL chain id
            for rnd in range(4, 8):
                whole.swap_step(T, rnd); a.swap_step(T, rnd); b.swap_step(T, rnd)
                assert np.array_equal(np.concatenate([a.state(), b.state()]), whole.state())
                whole.run(200); a.run(200); b.run(200)
        else:
            with pytest.raises(E.NativeError):
                a.swap_step(T, 99)          # a ladder of 16 does not fit a shard of 8


def test_cold_chains_sample_the_untempered_posterior(device):
    lin, ds = _dataset(device)
    T, ladders, n_iter, s = 4, 48, 120000, 100
    plain = E.Chains(ds, 64, seed=100)
    rp = plain.run(n_iter + 1, s)
    temp = E.Chains(ds, T * ladders, seed=200)
    temp.set_beta(np.tile(P.temperature_ladder(T, 0.15), ladders))
    rt, rounds = temp.run_tempered(n_iter + 1, s, T, 1000)
    assert rounds >= n_iter // 1000
    cold = E.cold_records(rt, T)
    assert cold.shape == (rp.shape[0], ladders, E.LR_REC_DOUBLES)
    hot = rt.reshape(rt.shape[0], ladders, T, -1)
    cnt = temp.counters()
    acc = cnt[:, 9].sum() / cnt[:, 8].sum()
    assert 0.05 < acc < 0.999

    def stats(r):
        b = r.shape[0] // 5
        return {k: r[b:, :, col].mean(0) for k, col in (("K_l", E.REC_KL), ("K_m", E.REC_KM), ("lik", E.REC_LIK), ("lam", E.REC_LAVG))}

    a, c = stats(rp), stats(cold)
    for k in a:
        se = np.sqrt(a[k].var(ddof=1) / len(a[k]) + c[k].var(ddof=1) / len(c[k]))
        z = abs(a[k].mean() - c[k].mean()) / se
        assert z < 4.5, (k, z, a[k].mean(), c[k].mean())
    # the hottest members see a flatter likelihood: lower mean log-likelihood than the cold ones
    hottest = np.take_along_axis(hot, np.argmin(hot[..., E.REC_BETA], axis=2)[:, :, None, None], axis=2)[:, :, 0, :]
    assert hottest[rt.shape[0] // 5:, :, E.REC_LIK].mean() < cold[rt.shape[0] // 5:, :, E.REC_LIK].mean()


def test_caller_stream_and_handle_stream_are_ordered(device):
    """run_device / swap_*_device default to torch's current stream, swap_step / state / counters / set_beta to the handle's
    own; the library orders the two (lr_order: an event recorded on the stream that touched the chains last, waited for by
    the next one) so that a caller who never synchronises gets what a caller who synchronises after every call gets."""
    lin, ds = _dataset(device)
    T, n = 8, 64
    beta = np.tile(P.temperature_ladder(T, 0.2), n // T)
    side = torch.cuda.Stream()

    def play(sync):
        ch = E.Chains(ds, n, seed=19)
        ch.set_beta(beta)
        for rnd in range(5):
            ch.run_device(30000, 0, None)                            # torch's current stream, asynchronous (milliseconds of work)
            if sync:
                torch.cuda.synchronize(); device.sync()
            ch.swap_step(T, rnd)                                     # the handle's stream
            if sync:
                torch.cuda.synchronize(); device.sync()
            with torch.cuda.stream(side):                            # a third stream
                ch.run_device(500, 0, None)
                info = ch.swap_info_device(torch.empty((n, 2), dtype=torch.float64, device="cuda"))
            if sync:
                torch.cuda.synchronize(); device.sync()
            ch.swap_apply_device(info, 0, T, rnd + 100)              # torch's current stream again, reading what `side` wrote
            if sync:
                torch.cuda.synchronize(); device.sync()
        return ch.state(), ch.counters()

    s0, c0 = play(True)
    s1, c1 = play(False)
    assert np.array_equal(s0, s1) and np.array_equal(c0, c1)
    assert c0[:, 9].sum() > 0
