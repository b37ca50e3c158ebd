"""GPU: the drop-in command line end to end (parse -> K1 -> K2/K3 -> the four log files)."""
import os

import numpy as np
import pytest

from conftest import GOLD, golden_input
from oracle import literate_oracle as O
from literate_b200 import forward as F

pytestmark = pytest.mark.gpu


def _read(path):
    with open(path) as fh:
        return fh.read().splitlines()


def _check_logs(stem, lin, st, model, n_rows, adequacy=True):
    mc = _read(stem + "_mcmc.log")
    assert mc[0].split("\t") == O.MCMC_HEADER + (O.ADEQUACY_HEADER if adequacy else [])
    sp, ex = _read(stem + "_sp_rates.log"), _read(stem + "_ex_rates.log")
    assert len(mc) - 1 == len(sp) == len(ex) == n_rows
    for i in (0, 1, n_rows // 2, n_rows - 1):
        row = mc[1 + i].split("\t")
        kl, km = int(row[6]), int(row[7])
        a, b = [float(x) for x in sp[i].split("\t")], [float(x) for x in ex[i].split("\t")]
        assert len(a) == 2 * kl - 1 and len(b) == 2 * km - 1          # rates then interior shift times, ragged
        L, M = np.array(a[:kl]), np.array(b[:km])
        tL = np.array([lin.start_time] + a[kl:] + [lin.end_time]); tM = np.array([lin.start_time] + b[km:] + [lin.end_time])
        lik = O.loglik_state(L, M, tL, tM, st, model)
        assert float(row[2]) == pytest.approx(lik, rel=1e-10)
        assert float(row[1]) == float(row[2]) + float(row[3])           # posterior = likelihood + prior (:321)
        assert float(row[4]) == pytest.approx(L.mean(), rel=1e-12) and float(row[5]) == pytest.approx(M.mean(), rel=1e-12)
        assert float(row[8]) == lin.start_time and float(row[9]) == lin.end_time


@pytest.mark.parametrize("model,suffix", [(0, "_BD"), (3, "_BDd")])
def test_single_chain_writes_the_reference_file_set(tmp_path, model, suffix, capsys):
    src = golden_input("example_dataTAD.txt", tmp_path)
    with pytest.warns(FutureWarning):
        F.main(["-d", src, "-n", "5001", "-s", "100", "-p", "2500", "-seed", "1", "-model_BDI", str(model)])
    out = capsys.readouterr().out
    assert "LiteRate" in out and "sp.times:" in out and "R^2:" in out
    d = os.path.join(str(tmp_path), "literate_mcmc_logs")
    assert sorted(os.listdir(d)) == sorted("example_dataTAD%s_%s.log" % (suffix, t) for t in ("div", "ex_rates", "mcmc", "sp_rates"))
    # div.log is deterministic: byte-identical to the file the unmodified reference wrote
    tag = "tad_m0" if model == 0 else "tad_m3"
    want = open(os.path.join(GOLD, "reference_logs", tag, "example_dataTAD%s_div.log" % suffix), "rb").read()
    assert open(os.path.join(d, "example_dataTAD%s_div.log" % suffix), "rb").read() == want
    lin = O.read_lineages(src)
    st = O.bin_stats(lin.ts, lin.te, only_dead=True, end_time=lin.end_time)
    _check_logs(os.path.join(d, "example_dataTAD" + suffix), lin, st, model, 51)


def test_many_chains_tbp_no_adequacy_and_tempering(tmp_path):
    src = golden_input("example_dataTBP.txt", tmp_path)
    with pytest.warns(FutureWarning):
        F.main(["-d", src, "-TBP", "-n", "3000", "-s", "50", "-seed", "3", "-chains", "5", "-calc_adequacy", "0", "-quiet", "1", "-out", "_t"])
    d = os.path.join(str(tmp_path), "literate_mcmc_logs")
    lin = O.read_lineages(src, TBP=True)
    st = O.bin_stats(lin.ts, lin.te)
    for k in range(5):
        stem = os.path.join(d, "example_dataTBP_BD_t_chain%d" % k)
        _check_logs(stem, lin, st, 0, 60, adequacy=False)
        assert os.path.exists(stem + "_div.log")
    rows = [open(os.path.join(d, "example_dataTBP_BD_t_chain%d_mcmc.log" % k)).read() for k in range(5)]
    assert len(set(rows)) == 5                      # independent chains
    # tempered: 2 logged chains, each the cold member of a ladder of 4
    with pytest.warns(FutureWarning):
        F.main(["-d", src, "-TBP", "-n", "4000", "-s", "100", "-seed", "3", "-chains", "2", "-temper", "4", "-swap_every", "200",
                "-quiet", "1", "-out", "_mc3"])
    for k in range(2):
        _check_logs(os.path.join(d, "example_dataTBP_BD_mc3_chain%d" % k), lin, st, 0, 40)


def test_logs_feed_the_posterior_summariser(tmp_path):
    """Four GPU chains -> log files -> combined posterior summary -> R script; the per-bin means agree
    with the same summary computed straight from the files by the oracle's restatement."""
    from literate_b200 import summary as S
    src = golden_input("example_dataTAD.txt", tmp_path)
    with pytest.warns(FutureWarning):
        F.main(["-d", src, "-n", "40001", "-s", "100", "-seed", "5", "-chains", "4", "-quiet", "1"])
    d = os.path.join(str(tmp_path), "literate_mcmc_logs")
    S.main([d, "-combine", "1", "-burnin", "0.25"])
    assert os.path.exists(os.path.join(d, "COMBINED_RTT_plots.r"))
    s = S.summarize_logs(os.path.join(d, "COMBINED_mcmc.log"), burnin=0, bf_seed=3)
    rows = [np.array(l.split(), float) for l in open(os.path.join(d, "COMBINED_sp_rates.log"))]
    m = O.marginal_rates(rows, s.death_age, s.root_age, 0)
    assert len(rows) == 4 * (401 - 100) and np.allclose(m.mean(0), s.birth.mean, rtol=1e-13)
    assert 0 < s.bf2 < s.bf6 < 1 and s.birth.k_counts.sum() == len(rows)


def test_directory_of_imputations(tmp_path):
    """-d <directory>: every table is a replicate; ragged tables are padded, each gets its own div.log and chains."""
    base = O.read_lineages(golden_input("example_dataTAD.txt"), death_jitter=0.0)
    rng = np.random.default_rng(8)
    d_in = tmp_path / "imps"
    d_in.mkdir()
    tabs = []
    for i in range(3):
        keep = rng.random(len(base.ts)) < (1.0 if i == 0 else 0.9)          # replicates of different length
        keep[[np.argmin(base.ts), np.argmax(base.te)]] = True                # same window for all
        ts, te = base.ts[keep], np.maximum(base.ts[keep], base.te[keep] - rng.integers(0, 2, keep.sum()) * (i > 0))
        te[np.argmax(base.te[keep])] = base.te.max()
        tabs.append((ts, te))
        with open(d_in / ("imp_%d.tsv" % i), "w") as fh:
            fh.write("id\tts\tte\n")
            for j, (a, b) in enumerate(zip(ts, te)):
                fh.write("%d\t%d\t%d\n" % (j, a, b))
    F.main(["-d", str(d_in), "-n", "2001", "-s", "100", "-seed", "4", "-chains", "6", "-quiet", "1"])
    d = os.path.join(str(d_in), "literate_mcmc_logs")
    for i, (ts, te) in enumerate(tabs):
        st = O.bin_stats(ts, te + 0.5)
        for j in range(2):
            stem = os.path.join(d, "imp_%d_BD_chain%d" % (i, j))
            div = np.loadtxt(stem + "_div.log", skiprows=1)
            assert np.array_equal(div[:, 0], st.sp) and np.array_equal(div[:, 1], st.ex) and np.array_equal(div[:, 2], st.br)
            lin = O.Lineages(ts, te + 0.5, base.ts.min(), base.te.max() + 0.5, 0)
            _check_logs(stem, lin, st, 0, 21)


def test_same_seed_same_files(tmp_path):
    a, b = tmp_path / "a", tmp_path / "b"
    a.mkdir(); b.mkdir()
    outs = []
    for sub in (a, b):
        src = golden_input("example_dataTAD.txt", sub)
        with pytest.warns(FutureWarning):
            F.main(["-d", src, "-n", "2000", "-s", "20", "-seed", "42", "-quiet", "1"])
        d = os.path.join(str(sub), "literate_mcmc_logs")
        outs.append({f: open(os.path.join(d, f), "rb").read() for f in sorted(os.listdir(d))})
    assert outs[0] == outs[1]


def test_long_multi_launch_run_loses_no_row(tmp_path, metal_path):
    """Several asynchronous launches of milliseconds each, 64 chains (the double-buffered loop of forward.run: the text of launch
    k is written while launch k + 1 runs): every chain's three logs are complete, and equal -- as text -- to the logs written
    from the records of ONE synchronous launch of the same chains."""
    import shutil
    from literate_b200 import engine as E
    src = os.path.join(str(tmp_path), "metal_bands_1.tsv")
    shutil.copy(metal_path, src)
    F.main(["-d", src, "-n", "300001", "-s", "1000", "-p", "100000000", "-seed", "5", "-chains", "64", "-quiet", "1", "-launch_iters", "40000"])
    d = os.path.join(str(tmp_path), "literate_mcmc_logs")
    lin = O.read_lineages(src)
    dev = E.Device(0)
    st = dev.bin_stats(lin.ts, lin.te)
    ds = E.Dataset(dev, st, 0, lin.start_time, lin.end_time)
    recs = E.Chains(ds, 64, 5).run(300001, 1000)                      # one launch, synchronous
    assert recs.shape[0] == 301
    for k in (0, 17, 63):
        stem = os.path.join(d, "metal_bands_1_BD_chain%d" % k)
        mc = _read(stem + "_mcmc.log")
        assert len(mc) == 302
        rows = np.array([[float(x) for x in l.split("\t")] for l in mc[1:]])
        assert np.array_equal(rows[:, 0], np.arange(301) * 1000.0)
        assert (rows[:, 6] >= 1).all() and (rows[:, 7] >= 1).all()          # no empty (never written) record
        assert np.array_equal(rows[:, 2], recs[:, k, E.REC_LIK]) and np.array_equal(rows[:, 6], recs[:, k, E.REC_KL])
