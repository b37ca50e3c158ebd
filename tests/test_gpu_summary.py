"""GPU: K5 (lr_summarize_records), posterior accumulators straight from device-resident sample records, against the host
summariser (literate_b200.summary, itself pinned to the unmodified plotRJforward.v3.py in tests/test_summary_host.py) and
the oracle's per-row restatement."""
import numpy as np
import pytest
import torch

from conftest import golden_input
from oracle import literate_oracle as O
from literate_b200 import engine as E, summary as S

pytestmark = pytest.mark.gpu


def test_device_accumulators_match_host_summary(device, metal_path):
    lin = O.read_lineages(metal_path)
    st = device.bin_stats(lin.ts, lin.te)
    ds = E.Dataset(device, st, 0, lin.start_time, lin.end_time)
    ch = E.Chains(ds, 48, seed=31)
    n_iter, s = 60001, 100
    rec = torch.empty((ch.records_per_run(n_iter, s), 48, E.LR_REC_DOUBLES), dtype=torch.float64, device="cuda")
    ch.run_device(n_iter, s, rec, stream="handle")
    device.sync()
    b0 = S.burnin_index(rec.shape[0], 0.2)
    post = rec[b0:].contiguous()
    nb = int(lin.end_time) - int(lin.start_time)
    sr, sc, kc, n = device.summarize_records_device(post, lin.start_time, nb)
    torch.cuda.synchronize()
    sr, sc, kc = sr.cpu().numpy(), sc.cpu().numpy(), kc.cpu().numpy()
    host = S.summarize_records(rec.cpu().numpy(), lin.start_time, lin.end_time, burnin=0.2, bf_seed=None)
    assert n == host.birth.n_samples == post.shape[0] * 48
    np.testing.assert_allclose(sr[0] / n, host.birth.mean, rtol=1e-12)
    np.testing.assert_allclose(sr[1] / n, host.death.mean, rtol=1e-12)
    assert np.array_equal(sc[0] / n, host.birth.shift_freq) and np.array_equal(sc[1] / n, host.death.shift_freq)
    for side, hs in ((0, host.birth), (1, host.death)):
        want = np.zeros(32, dtype=np.int64)
        want[hs.k_values.astype(int) - 1] = hs.k_counts
        assert np.array_equal(kc[side], want)
    # oracle restatement on a few hundred rows
    r = rec.cpu().numpy()[b0:, 0]
    rows = [np.concatenate([x[E.REC_L:E.REC_L + int(x[E.REC_KL])], x[E.REC_TL + 1:E.REC_TL + int(x[E.REC_KL])]]) for x in r]
    sr1, _, _, n1 = device.summarize_records_device(post[:, 0].contiguous(), lin.start_time, nb)
    np.testing.assert_allclose(sr1[0].cpu().numpy() / n1, O.marginal_rates(rows, lin.end_time, lin.start_time, 0).mean(0), rtol=1e-12)
    # the summary module's device entry point: same numbers, burn-in handled there
    dv = S.summarize_records_device(device, rec, lin.start_time, lin.end_time, burnin=0.2)
    np.testing.assert_allclose(dv["birth"]["mean"], host.birth.mean, rtol=1e-12)
    np.testing.assert_allclose(dv["net_mean"], host.net_mean, rtol=1e-10, atol=1e-15)
    assert np.array_equal(dv["death"]["k_values"], host.death.k_values) and np.array_equal(dv["death"]["k_counts"], host.death.k_counts)
    assert np.array_equal(dv["time"], host.birth.time)
    # deterministic
    sr2, sc2, kc2, _ = device.summarize_records_device(post, lin.start_time, nb)
    assert torch.equal(sr2.cpu(), torch.from_numpy(sr)) and torch.equal(sc2.cpu(), torch.from_numpy(sc))


def test_more_than_256_bins_and_edge_shifts(device):
    """Hand-made records: 300 bins (two passes), shift times on bin edges, on the closed last edge and outside the edges."""
    nb, e0 = 300, 100.0
    rec = np.zeros((5, E.LR_REC_DOUBLES))
    cases = [([0.5], []), ([0.1, 0.2], [150.0]), ([1., 2., 3.], [100.0, 400.0]), ([1., 2., 3., 4.], [99.5, 250.25, 401.0]), ([5., 6.], [399.999])]
    for r, (rates, shifts) in zip(rec, cases):
        k = len(rates)
        r[E.REC_KL], r[E.REC_KM] = k, 1
        r[E.REC_L:E.REC_L + k] = rates; r[E.REC_TL + 1:E.REC_TL + k] = shifts
        r[E.REC_M] = 0.25
    sr, sc, kc, n = device.summarize_records_device(torch.from_numpy(rec).cuda(), e0, nb)
    torch.cuda.synchronize()
    edges = np.arange(e0, e0 + nb + 1)
    want = np.zeros(nb); cnt = np.zeros(nb, dtype=np.int64)
    for rates, shifts in cases:
        h = np.histogram(shifts, bins=edges)[0]
        want += np.array(rates)[np.cumsum(h)]; cnt += h
    assert np.array_equal(sr[0].cpu().numpy(), want) and np.array_equal(sc[0].cpu().numpy(), cnt)
    assert np.array_equal(sr[1].cpu().numpy(), np.full(nb, 5 * 0.25))
    assert kc[0].cpu().numpy()[:4].tolist() == [1, 2, 1, 1] and kc[1].cpu().numpy()[0] == 5


def test_hpd_intervals_on_the_device(device, metal_path):
    """The per-sample matrix (lr_marginal_rates) equals the host's marginal_matrix element for element, and the HPD intervals
    of birth, death and net rate computed on the device (sorted there, a few bins at a time) are the host's, bit for bit."""
    lin = O.read_lineages(metal_path)
    st = device.bin_stats(lin.ts, lin.te)
    ds = E.Dataset(device, st, 0, lin.start_time, lin.end_time)
    ch = E.Chains(ds, 24, seed=7)
    n_iter, s = 40001, 100
    rec = torch.empty((ch.records_per_run(n_iter, s), 24, E.LR_REC_DOUBLES), dtype=torch.float64, device="cuda")
    ch.run_device(n_iter, s, rec, stream="handle")
    device.sync()
    host = S.summarize_records(rec.cpu().numpy(), lin.start_time, lin.end_time, burnin=0.2, bf_seed=None)
    b0 = S.burnin_index(rec.shape[0], 0.2)
    post = rec[b0:].contiguous()
    nb = int(lin.end_time) - int(lin.start_time)
    mb, md = device.marginal_rates_device(post, lin.start_time, nb)
    torch.cuda.synchronize()
    assert np.array_equal(mb.cpu().numpy(), host.birth.marginal) and np.array_equal(md.cpu().numpy(), host.death.marginal)
    mb2, md2 = device.marginal_rates_device(post, lin.start_time, nb, 5, 9)          # a range of bins
    assert torch.equal(mb2, mb[:, 5:14]) and torch.equal(md2, md[:, 5:14])
    dv = S.summarize_records_device(device, rec, lin.start_time, lin.end_time, burnin=0.2, hpd=True, hpd_bytes=8 * post.shape[0] * post.shape[1] * 7)
    for name, hs in (("birth", host.birth), ("death", host.death)):
        assert np.array_equal(dv[name]["hpd_lo"], hs.hpd_lo) and np.array_equal(dv[name]["hpd_hi"], hs.hpd_hi)
    assert np.array_equal(dv["net_lo"], host.net_lo) and np.array_equal(dv["net_hi"], host.net_hi)


def test_imputation_envelope_on_device_equals_the_host_restatement(device):
    """lr_imputation_envelope on K1's device output (replicate axis = imputations) against summary.imputation_envelope, which
    tests/test_summary_host.py pins to the unmodified utilities/imputation_averager.py: bit for bit, NaN bins included."""
    import torch
    from literate_b200 import summary as S, synth
    tdev = torch.device("cuda:0")
    n_rep, n, nb = 7, 20000, 200
    ts, te = synth.syn_int_device(n, n_rep, tdev)
    ts[:, 1:n] = torch.clamp(ts[:, 1:n], min=1803.0)          # bins 1 and 2 see no births, bin 0 exactly one lineage
    sp, ex, br = device.bin_stats_device(ts[:, :n], te[:, :n], 1800, nb)
    br[2, 5] = 0.0; sp[2, 5] = 0; ex[2, 5] = 0                 # an empty bin in one replicate: 0/0 = NaN, which numpy's min/max propagate
    out = device.imputation_envelope_device(sp, ex, br).cpu().numpy()
    tables = np.stack([sp.cpu().numpy().astype(float), ex.cpu().numpy().astype(float), br.cpu().numpy()], axis=2)
    env = S.imputation_envelope(tables)
    for k, name in enumerate(S.ENVELOPE_ROWS):
        assert np.array_equal(out[k], env[name], equal_nan=True), name
    assert np.isnan(out[:6, 5]).all()
