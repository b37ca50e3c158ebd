"""CPU: the host side of the drop-in boundary (no kernel is called here).

* the command line accepts the reference's 21 flags with the reference's defaults (LiteRateForward.py:376-401);
* input parsing equals the oracle's restatement of :439-476 on the shipped tables;
* the log writers reproduce the reference's own files byte for byte when fed the numbers those files hold
  (row layout, ragged sp/ex rows, shortest-repr floats, the '1' of the untouched Poisson rate, -pyrate_output,
  the \\r\\n rows of div.log);
* the C-ABI library loads and exports every symbol include/literate_b200.h declares.
"""
import ctypes
import json
import os
import re

import numpy as np
import pytest

from conftest import GOLD, REPO, golden_input
from oracle import literate_oracle as O
from literate_b200 import _native, engine as E, forward as F

REFERENCE_DEFAULTS = {   # LiteRateForward.py:378-401
    "d": "", "n": 10000000, "p": 1000, "s": 1000, "seed": -1, "const_rates": 0, "const_death_rate": 0, "model_BDI": 0,
    "TBP": False, "pyrate_output": False, "first_year": -1, "last_year": -1, "death_jitter": .5, "use_rate_HP": 1,
    "Poisson_prior": 0, "rm_first_bin": 0, "calc_adequacy": 1, "update_fraction": 0.75, "out": "", "rev_se": 0,
}


def test_parser_has_the_reference_flags_and_defaults():
    a = F.build_parser().parse_args([])
    for k, v in REFERENCE_DEFAULTS.items():
        assert getattr(a, k) == v, k
    assert a.chains == 1 and a.real_move_shift == 0          # new flags default to the reference's behaviour
    a = F.build_parser().parse_args("-d x.tsv -n 5 -s 2 -p 3 -seed 9 -model_BDI 3 -TBP -pyrate_output -death_jitter 0 "
                                    "-Poisson_prior 2.5 -update_fraction .5 -out _x -rev_se 1 -const_death_rate 1".split())
    assert (a.d, a.n, a.s, a.p, a.seed, a.model_BDI, a.TBP, a.pyrate_output) == ("x.tsv", 5, 2, 3, 9, 3, True, True)
    assert (a.death_jitter, a.Poisson_prior, a.update_fraction, a.out, a.rev_se, a.const_death_rate) == (0.0, 2.5, .5, "_x", 1, 1)


@pytest.mark.parametrize("name,tbp", [("example_dataTAD.txt", False), ("example_dataTBP.txt", True)])
def test_parse_lineages_equals_oracle(name, tbp):
    for jitter in (0.5, 0.0):
        with pytest.warns(FutureWarning):        # 4-column layout is deprecated (:441-444)
            ts, te, s, e, root = F.parse_lineages(golden_input(name), TBP=tbp, death_jitter=jitter)
        lin = O.read_lineages(golden_input(name), TBP=tbp, death_jitter=jitter)
        assert np.array_equal(ts, lin.ts) and np.array_equal(te, lin.te)
        assert (s, e, root) == (lin.start_time, lin.end_time, lin.true_root_age)


def test_parse_three_columns_rev_se_and_year_filters(tmp_path):
    p = tmp_path / "t.tsv"
    p.write_text("id\tts\tte\r\n1\t1990\t1995\t\r\n2\t1985\t2001\r\n3\t1999\t2003\r\n")
    ts, te, s, e, _ = F.parse_lineages(str(p))
    assert ts.tolist() == [1990, 1985, 1999] and te.tolist() == [1995.5, 2001.5, 2003.5] and (s, e) == (1985, 2003.5)
    ts2, te2, *_ = F.parse_lineages(str(p), rev_se=1)
    assert ts2.tolist() == [1995, 2001, 2003] and te2.tolist() == [1990.5, 1985.5, 1999.5]
    ts3, te3, *_ = F.parse_lineages(str(p), first_year=1988, last_year=1995)      # intended semantics (literate_library.py:216-222)
    assert ts3.tolist() == [1990] and te3.tolist() == [1995.5]
    lin = O.read_lineages(str(p), first_year=1988, last_year=1995)
    assert np.array_equal(lin.ts, ts3) and np.array_equal(lin.te, te3)


def _jobs():
    with open(os.path.join(GOLD, "reference_logs", "manifest.json")) as fh:
        return json.load(fh)


def _records_from_logs(mcmc, sp, ex, pyrate, root, start, end):
    """The sample records a device run would have delivered for these reference rows."""
    rows = [l.split("\t") for l in mcmc.splitlines()[1:]]
    recs = np.zeros((len(rows), E.LR_REC_DOUBLES))
    for i, (r, ls, le) in enumerate(zip(rows, sp.splitlines(), ex.splitlines())):
        kl, km = int(r[6]), int(r[7])
        rec = recs[i]
        rec[E.REC_IT], rec[E.REC_LIK], rec[E.REC_PRIOR] = float(r[0]), float(r[2]), float(r[3])
        rec[E.REC_LAVG], rec[E.REC_MAVG], rec[E.REC_KL], rec[E.REC_KM] = float(r[4]), float(r[5]), kl, km
        rec[E.REC_GL], rec[E.REC_GM] = float(r[10]), float(r[11])
        rec[E.REC_POI_INIT] = 1.0 if re.fullmatch(r"\d+", r[12]) else 0.0       # a Python int: never touched by the Gibbs step
        rec[E.REC_POI] = float(r[12])
        if len(r) > 13:
            rec[E.REC_ADQ:E.REC_ADQ + 3] = [float(x) for x in r[13:16]]
        fl, fe = [float(x) for x in ls.split("\t")], [float(x) for x in le.split("\t")]
        assert len(fl) == 2 * kl - 1 and len(fe) == 2 * km - 1
        rec[E.REC_L:E.REC_L + kl], rec[E.REC_M:E.REC_M + km] = fl[:kl], fe[:km]
        sl, se = np.array(fl[kl:]), np.array(fe[km:])
        if pyrate:
            sl, se = root - sl, root - se
        rec[E.REC_TL], rec[E.REC_TM] = start, start
        rec[E.REC_TL + 1:E.REC_TL + kl], rec[E.REC_TM + 1:E.REC_TM + km] = sl, se
    return recs


@pytest.mark.parametrize("job", _jobs(), ids=lambda j: j["tag"])
def test_log_writers_reproduce_reference_files(job, tmp_path):
    a = F.build_parser().parse_args(["-d", job["input"]] + job["args"])
    d = os.path.join(GOLD, "reference_logs", job["tag"])
    files = {f.rsplit("_", 1)[-1] if not f.endswith("rates.log") else "_".join(f.rsplit("_", 2)[-2:]): f for f in job["files"]}
    want = {k: open(os.path.join(d, f), "rb").read() for k, f in files.items()}
    lin = O.read_lineages(golden_input(job["input"], tmp_path), TBP=a.TBP, death_jitter=a.death_jitter)
    pyrate_shifts_exact = True
    recs = _records_from_logs(want["mcmc.log"].decode(), want["sp_rates.log"].decode(), want["ex_rates.log"].decode(),
                              a.pyrate_output, lin.true_root_age, lin.start_time, lin.end_time)
    stem = os.path.join(str(tmp_path), "out")
    w = F.ChainLogWriter(stem, a.calc_adequacy, a.pyrate_output, lin.start_time, lin.end_time, lin.true_root_age, a.Poisson_prior)
    for r in recs:
        w.write(r)
    w.close()
    assert open(stem + "_mcmc.log", "rb").read() == want["mcmc.log"]
    if not a.pyrate_output:
        assert open(stem + "_sp_rates.log", "rb").read() == want["sp_rates.log"]
        assert open(stem + "_ex_rates.log", "rb").read() == want["ex_rates.log"]
    else:
        # root - (root - t) is not always t to the last bit: compare numerically
        for tag in ("sp_rates.log", "ex_rates.log"):
            got = [[float(x) for x in l.split("\t")] for l in open(stem + "_" + tag).read().splitlines()]
            ref = [[float(x) for x in l.split("\t")] for l in want[tag].decode().splitlines()]
            assert len(got) == len(ref)
            for g, r in zip(got, ref):
                np.testing.assert_allclose(g, r, rtol=0, atol=1e-9)
    # div.log from the oracle's statistics through the product's writer
    st = O.bin_stats(lin.ts, lin.te)
    F.write_div_log(stem + "_div.log", st.sp, st.ex, st.br)
    assert open(stem + "_div.log", "rb").read() == want["div.log"]


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(REPO, "include", "literate_b200.h")).read()
    declared = set(re.findall(r"\b(lr_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_native.EXPORTS), declared ^ set(_native.EXPORTS)
    from literate_b200 import build
    lib = ctypes.CDLL(build.build())
    for s in declared:
        assert hasattr(lib, s), s
    assert lib.lr_abi_version() == _native.LR_ABI_VERSION
    # without a GPU the library refuses to create a handle and says why -- it does not fall back
    import torch
    if not torch.cuda.is_available():
        n = _native.load()
        h = ctypes.c_void_p()
        assert n.lr_create(0, ctypes.byref(h)) != 0
        assert b"no CPU path" in n.lr_last_error()


def test_reference_cost_binning_equals_oracle():
    rng = np.random.default_rng(4)
    ts = 100 + np.floor(rng.uniform(0, 30, 500)); te = ts + np.floor(rng.exponential(5, 500)) + .5
    for j in range(100, 131):
        assert O.events_in_bin_asref(ts, te, j, j + 1) == O.events_in_bin(ts, te, j, j + 1)


def test_integer_year_tables_are_recognised():
    """forward.as_year_table: the tables the reference ships become int32 year tables (the compact K1 entry point), anything
    with a fractional time, an out-of-range jitter or a lone NaN does not."""
    from literate_b200 import engine as E
    ts, te, _, _, _ = F.parse_lineages(golden_input("example_dataTAD.txt"))
    ti, ei = F.as_year_table(ts, te, 0.5)
    assert ti.dtype == np.int32 and np.array_equal(ti, ts) and np.array_equal(ei + 0.5, te)
    assert F.as_year_table(ts + 0.25, te, 0.5) is None and F.as_year_table(ts, te + 0.1, 0.5) is None and F.as_year_table(ts, te, 1.5) is None
    pad = np.concatenate([ts, [np.nan, np.nan]]), np.concatenate([te, [np.nan, np.nan]])      # a ragged replicate's NaN rows
    ti, ei = F.as_year_table(*pad, 0.5)
    assert (ti[-2:] == E.YEAR_PAD).all() and (ei[-2:] == E.YEAR_PAD).all() and np.array_equal(ti[:-2], ts)
    assert F.as_year_table(np.concatenate([ts, [np.nan]]), np.concatenate([te, [2000.5]]), 0.5) is None
