"""CPU: the TrendRate oracle against the logs the UNMODIFIED trend_rate.py wrote in the build container
(tests/golden/trendrate, made by oracle/make_golden_trend.py).  With the same seed the oracle chain must reproduce each
log byte for byte: parsing, bins (last one dropped), trend normalisation, likelihood, priors, both proposal kinds, the
strict accept rule, the forced acceptance of iteration 0, adequacy and the csv formatting."""
import gzip
import json
import os
import shutil

import numpy as np
import pytest

from conftest import GOLD
from oracle import trendrate_oracle as T

TG = os.path.join(GOLD, "trendrate")


def _flag(args, name, default, cast=float):
    return cast(args[args.index(name) + 1]) if name in args else default


def _jobs():
    with open(os.path.join(TG, "manifest.json")) as fh:
        return json.load(fh)


def stage(job, tmp_path):
    """Copies the job's inputs into tmp_path; returns (data path, trend path)."""
    for name in job["inputs"]:
        src = os.path.join(TG, name)
        if not os.path.exists(src):
            src = os.path.join(GOLD, "inputs", name)
        dst = os.path.join(str(tmp_path), name.replace(".gz", ""))
        if src.endswith(".gz"):
            with gzip.open(src, "rb") as a, open(dst, "wb") as b:
                shutil.copyfileobj(a, b)
        else:
            shutil.copy(src, dst)
    a = job["args"]
    return os.path.join(str(tmp_path), job["data"]), os.path.join(str(tmp_path), a[a.index("-trend_data") + 1])


def setup_job(job, tmp_path):
    a = job["args"]
    data, trend_path = stage(job, tmp_path)
    rm = _flag(a, "-rm_first_bin", 0.0)
    ts, te, present, origin = T.parse_ts_te(data, death_jitter=_flag(a, "-death_jitter", 0.5))
    bins = T.create_bins(origin, present, ts, te, rm)
    idx = _flag(a, "-trend_index", 0, int)
    trend = T.normalise_trend(T.read_trend_column(trend_path, idx), rm)
    flags = dict(const_birth="-const_B" in a, const_death="-const_D" in a, no_death="-no_death" in a)
    return data, bins, trend, idx, flags


@pytest.mark.parametrize("job", _jobs(), ids=lambda j: j["tag"])
def test_oracle_chain_reproduces_reference_log(job, tmp_path):
    a = job["args"]
    data, bins, trend, idx, flags = setup_job(job, tmp_path)
    seed = _flag(a, "-seed", -1, int)
    name = T.log_name(data, seed, idx, **flags)
    with open(name, "w", newline="") as fh:
        T.run_chain(bins, trend, _flag(a, "-n", 0, int), _flag(a, "-s", 1000, int), seed, fh, exact_scipy=True, **flags)
    (f,) = job["files"]
    assert os.path.basename(name) == f
    want = open(os.path.join(TG, job["tag"], f), "rb").read()
    got = open(name, "rb").read()
    assert got == want, f"{job['tag']}/{f} differs from the reference's own output"


def test_closed_form_priors_match_scipy():
    rng = np.random.default_rng(3)
    for _ in range(50):
        p = np.array([rng.gamma(1, .3) + .002, rng.gamma(1, .3) + .002, rng.normal(0, 2), rng.normal(0, 2),
                      rng.gamma(3, .5), rng.gamma(3, .5)])
        assert abs(T.prior(p) - T.prior(p, exact_scipy=True)) < 1e-12
    assert T.prior(np.array([.0005, .1, 0, 0, 1, 1])) == -np.inf       # below the location .001 of the Gamma prior
    assert T.prior(np.array([.1, .1, 0, 0, -1, 1])) == -np.inf


def test_bins_drop_the_last_interval():
    job = [j for j in _jobs() if j["tag"] == "ex_ramp"][0]
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        _, bins, trend, _, _ = setup_job(job, d)
    assert bins.origin == 1994 and bins.n_bins == 24 and len(trend) == 24
    assert trend.min() == T.SMALL_NUMBER and trend.max() == 1.0
    assert int(bins.n_spec.sum()) == 75
