"""CPU: the DDRate oracle against the logs the UNMODIFIED DDRatev3.py wrote in the build container
(tests/golden/ddrate, made by oracle/make_golden_ddrate.py; only -m_birth 3 with a genre table runs as shipped -- the
-m_birth 0 / 1 / 2 fixtures come from the script's unmodified text executed with the three names its line 48 lacks pre-bound).
With the same seed the oracle chain must reproduce the sample log and the div.log byte for byte."""
import gzip
import json
import os
import shutil

import numpy as np
import pytest

from conftest import GOLD
from oracle import ddrate_oracle as D

DG = os.path.join(GOLD, "ddrate")


def _flag(args, name, default, cast=float):
    return cast(args[args.index(name) + 1]) if name in args else default


def _jobs():
    with open(os.path.join(DG, "manifest.json")) as fh:
        return json.load(fh)


def stage(job, tmp_path):
    """Copies the job's inputs into tmp_path; returns (data path, genre path)."""
    for name in job["inputs"]:
        src = [p for p in (os.path.join(DG, name), os.path.join(GOLD, "trendrate", name), os.path.join(GOLD, "inputs", name)) if os.path.exists(p)][0]
        dst = os.path.join(str(tmp_path), name.replace(".gz", ""))
        if src.endswith(".gz"):
            with gzip.open(src, "rb") as a, open(dst, "wb") as b:
                shutil.copyfileobj(a, b)
        else:
            shutil.copy(src, dst)
    a = job["args"]
    return os.path.join(str(tmp_path), job["data"]), (os.path.join(str(tmp_path), a[a.index("-g") + 1]) if "-g" in a else None)


def setup_job(job, tmp_path):
    a = job["args"]
    data, genre = stage(job, tmp_path)
    rm = _flag(a, "-rm_first_bin", 0.0)
    ts, te, present, origin = D.parse_ts_te(data)
    bins = D.create_bins(origin, present, ts, te, rm)
    if genre is None:
        # -m_birth 0 / 1 / 2: no genre table.  The shipped script cannot start these (NameError at :48); the fixtures were written
        # by its unmodified text run with the three missing names pre-bound (oracle/make_golden_ddrate.py, SHIM_JOBS)
        return data, D.Setup(bins, _flag(a, "-m_birth", 2, int), _flag(a, "-m_death", 2, int), None, None), None
    gts, gte, gpresent, gorigin = D.parse_ts_te(genre)
    gbins = D.create_bins(gorigin, gpresent, gts, gte, rm)
    S = D.Setup(bins, _flag(a, "-m_birth", 2, int), _flag(a, "-m_death", 2, int), gts, gte)
    return data, S, gbins


@pytest.mark.parametrize("job", _jobs(), ids=lambda j: j["tag"])
def test_oracle_chain_reproduces_reference_logs(job, tmp_path):
    a = job["args"]
    data, S, gbins = setup_job(job, tmp_path)
    seed = _flag(a, "-seed", -1, int)
    stem = D.log_stem(data, seed, S.m_birth, S.m_death)
    with open(stem + ".log", "w", newline="") as fh:
        D.run_chain(S, _flag(a, "-n", 0, int), _flag(a, "-s", 1000, int), seed, fh, exact_scipy=True)
    with open(stem + ".div.log", "w", newline="") as fh:
        D.write_div_log(fh, S.bins, gbins)
    assert sorted(os.path.basename(stem) + e for e in (".log", ".div.log")) == job["files"]
    for f in job["files"]:
        want = open(os.path.join(DG, job["tag"], f), "rb").read()
        got = open(os.path.join(str(tmp_path), f), "rb").read()
        assert got == want, f"{job['tag']}/{f} differs from the reference's own output"


def test_closed_form_priors_match_scipy(tmp_path):
    _, S, _ = setup_job(_jobs()[0], tmp_path)
    rng = np.random.default_rng(5)
    for _ in range(50):
        p = np.array([rng.gamma(1, .5), rng.gamma(1, 1), rng.gamma(1, 2), rng.uniform(0, 20), rng.gamma(1, 10), rng.gamma(1, 50),
                      rng.uniform(0, 1), rng.gamma(3, .5), rng.gamma(3, .5), rng.gamma(1, 1), rng.gamma(1, 1)])
        assert abs(D.prior(p, S) - D.prior(p, S, exact_scipy=True)) < 1e-12
    p[3] = S.present - S.origin                       # midpoint at the present: rejected (:139-140)
    assert D.prior(p, S) == -np.inf
    p[3] = 3.0; p[6] = 1.0
    assert D.prior(p, S) == -np.inf and D.prior(p, S, exact_scipy=True) == -np.inf
