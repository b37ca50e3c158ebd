"""CPU: posterior summaries against what the UNMODIFIED plotRJforward.v3.py wrote for the same log
files (tests/golden/plots/*.r, made by `oracle/make_golden.py plots`) and against the oracle's restatement."""
import os
import re

import numpy as np
import pytest

from conftest import GOLD
from oracle import literate_oracle as O
from literate_b200 import engine as E, summary as S


def _r_vectors(path):
    """name -> list of floats for every `name=c(...)` / `name=as.numeric(c(...))` line of an R script (in order of appearance)."""
    out = {}
    for line in open(path):
        m = re.match(r"^([A-Za-z0-9_]+)=(?:as\.numeric\()?c\((.*?)\)+\s*$", line)
        if m:
            vals = [np.nan if x.strip() in ("NA", "nan", "") else float(x) for x in m.group(2).split(",")] if m.group(2).strip() else []
            out.setdefault(m.group(1), []).append(vals)
    return out


@pytest.mark.parametrize("tag,stem", [("tad_m0", "example_dataTAD_BD"), ("metal_m0", "metal_bands_1_BD")])
def test_summary_equals_reference_plotter(tag, stem, tmp_path):
    want = _r_vectors(os.path.join(GOLD, "plots", tag + "_RTT_plots.r"))
    s = S.summarize_logs(os.path.join(GOLD, "reference_logs", tag, stem + "_mcmc.log"), burnin=0.2, bf_seed=1)
    eq = lambda a, b: np.testing.assert_allclose(np.asarray(a, float), np.asarray(b, float), rtol=1e-13, atol=0, equal_nan=True)
    eq(s.birth.k_values, want["unique"][0]); eq(s.birth.k_counts, want["counts"][0])
    eq(s.death.k_values, want["unique"][1]); eq(s.death.k_counts, want["counts"][1])
    eq(s.birth.time, want["time"][0])
    for side, p in ((s.birth, "birth"), (s.death, "death")):
        eq(side.mean, want[p + "_rate"][0]); eq(side.hpd_lo, want[p + "_minHPD"][0]); eq(side.hpd_hi, want[p + "_maxHPD"][0])
        eq(side.shift_freq, want[p + "_counts"][0])
    eq(s.net_mean, want["net_rate"][0]); eq(s.net_lo, want["net_minHPD"][0]); eq(s.net_hi, want["net_maxHPD"][0])
    eq(s.div[:, 2], want["net_diversity"][0])
    # the Bayes-factor thresholds come from a prior simulation (unseeded in the reference): Monte-Carlo agreement
    ref_bf = [float(l.split("=")[1]) for l in open(os.path.join(GOLD, "plots", tag + "_RTT_plots.r")) if l.startswith("bf")]
    assert s.bf2 == pytest.approx(ref_bf[0], rel=0.03) and s.bf6 == pytest.approx(ref_bf[1], rel=0.03)
    # our own R file defines the same variables with the same numbers
    mine = _r_vectors(S.write_r([s], str(tmp_path / "x_RTT_plots.r")))
    for k in ("unique", "counts", "time", "birth_rate", "birth_minHPD", "birth_maxHPD", "birth_counts", "death_rate", "death_minHPD",
              "death_maxHPD", "death_counts", "net_rate", "net_minHPD", "net_maxHPD", "net_diversity"):
        assert len(mine[k]) >= 1
        eq(mine[k][0], want[k][0])


def test_records_path_equals_log_path_and_oracle():
    """The same numbers from sample records (no text round trip) and from the oracle's per-row restatement."""
    d = os.path.join(GOLD, "reference_logs", "metal_m0")
    mc = os.path.join(d, "metal_bands_1_BD_mcmc.log")
    s_log = S.summarize_logs(mc, bf_seed=None)
    tbl = np.loadtxt(mc, skiprows=1)
    rows_l = [np.array(l.split(), float) for l in open(mc.replace("mcmc.log", "sp_rates.log"))]
    rows_m = [np.array(l.split(), float) for l in open(mc.replace("mcmc.log", "ex_rates.log"))]
    rec = np.zeros((len(tbl), E.LR_REC_DOUBLES))
    for i, (t, a, b) in enumerate(zip(tbl, rows_l, rows_m)):
        kl, km = int(t[6]), int(t[7])
        rec[i, E.REC_KL], rec[i, E.REC_KM] = kl, km
        rec[i, E.REC_L:E.REC_L + kl], rec[i, E.REC_TL + 1:E.REC_TL + kl] = a[:kl], a[kl:]
        rec[i, E.REC_M:E.REC_M + km], rec[i, E.REC_TM + 1:E.REC_TM + km] = b[:km], b[km:]
    s_rec = S.summarize_records(rec, s_log.root_age, s_log.death_age, bf_seed=None)
    for a, b in ((s_rec.birth, s_log.birth), (s_rec.death, s_log.death)):
        assert np.array_equal(a.mean, b.mean) and np.array_equal(a.hpd_lo, b.hpd_lo) and np.array_equal(a.hpd_hi, b.hpd_hi)
        assert np.array_equal(a.shift_freq, b.shift_freq) and np.array_equal(a.k_counts, b.k_counts)
    assert np.array_equal(s_rec.net_mean, s_log.net_mean)
    # oracle restatement of get_marginal_rates, row by row
    m = O.marginal_rates(rows_l, s_log.death_age, s_log.root_age, 0.2)
    assert np.allclose(m.mean(0), s_log.birth.mean, rtol=1e-14)
    assert O.k_pmf(tbl[:, 6]) == dict(zip(s_log.birth.k_values.astype(int).tolist(), s_log.birth.k_counts.tolist()))


def test_hpd_matches_the_scalar_definition():
    rng = np.random.default_rng(0)
    x = rng.gamma(2.0, 1.0, (57, 5))
    lo, hi = S.hpd_columns(x)
    for j in range(5):
        d = np.sort(x[:, j]); n_in = int(round(0.95 * len(d)))
        w = [d[k + n_in - 1] - d[k] for k in range(len(d) - n_in + 1)]
        k = int(np.argmin(w))
        assert (lo[j], hi[j]) == (d[k], d[k + n_in - 1])


def test_combine_logs_and_cli(tmp_path):
    import shutil
    for tag in ("tad_m0", "tad_m2"):
        for f in os.listdir(os.path.join(GOLD, "reference_logs", tag)):
            shutil.copy(os.path.join(GOLD, "reference_logs", tag, f), tmp_path)
    files = sorted(str(p) for p in tmp_path.glob("*mcmc.log"))
    assert len(files) == 2
    out = S.combine_logs(files, str(tmp_path), burnin=0.25)
    rows = open(out).read().splitlines()
    n_each = [len(open(f).read().splitlines()) - 1 for f in files]
    assert len(rows) - 1 == sum(n - int(0.25 * n) for n in n_each)
    assert [r.split("\t")[0] for r in rows[1:4]] == ["0", "1", "2"]                       # `it` renumbered
    assert len(open(out.replace("mcmc.log", "sp_rates.log")).read().splitlines()) == len(rows) - 1
    div = np.loadtxt(out.replace("mcmc.log", "div.log"), skiprows=1)
    assert div.shape == (24, 3)
    S.main([str(tmp_path), "-combine", "1", "-burnin", "0.25"])
    assert os.path.exists(os.path.join(str(tmp_path), "COMBINED_RTT_plots.r"))


def test_imputation_averager_prints_what_the_reference_utility_prints(tmp_path, capsys):
    """utilities/imputation_averager.py:23-61 (mean / min / max over imputation replicates of sp/br, ex/br and br): the fixture
    holds five div.log files and the stdout of the UNMODIFIED utility on them (oracle/make_golden.py averager).  Same text,
    byte for byte, when the files are read in the order the utility read them; same numbers to 1e-15 in any order."""
    import json
    import shutil
    from literate_b200 import summary as S
    src = os.path.join(GOLD, "averager")
    want = open(os.path.join(src, "expected_stdout.txt")).read()
    order = json.load(open(os.path.join(src, "file_order.json")))
    tables = np.stack([np.loadtxt(os.path.join(src, f), skiprows=1, ndmin=2) for f in order])
    env = S.imputation_envelope(tables)
    assert S.format_imputation_envelope(env) == want
    # the command-line path (glob order of THIS file system: means may differ in the last bit)
    for f in order:
        shutil.copy(os.path.join(src, f), tmp_path)
    S.main([str(tmp_path), "-imputations", "1"])
    got = capsys.readouterr().out
    assert got.splitlines()[0] == "EMPIRICAL DEATH" and len(got.splitlines()) == len(want.splitlines())

    def vec(line):
        return np.array([float(x) if x.strip() != "NA" else np.nan for x in line[line.index("(") + 1:line.rindex(")")].split(",")])
    for a, b in zip(got.splitlines(), want.splitlines()):
        if "=c(" in a:
            np.testing.assert_allclose(vec(a), vec(b), rtol=1e-15, equal_nan=True)
    assert np.isnan(env["ed_mean"]).sum() == 0 and env["eb_max"][0] == 2.0      # the half-year bin of replicate 3: one birth / 0.5
