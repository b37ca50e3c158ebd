#!/usr/bin/env python
"""Generate tests/golden/trendrate/ by running the UNMODIFIED trend_rate.py in this container.

TEST INFRASTRUCTURE ONLY.  Needs /root/reference (read-only, absent on the GPU box); everything it produces is committed.

    python oracle/make_golden_trend.py kat         # seconds: exact logs of short reference runs
    python oracle/make_golden_trend.py posterior   # ~4 min: 8 reference chains per configuration -> posterior summaries

The reference tree ships no trend table (trend_rate.py needs ``-trend_data``) and literate_library.parse_ts_te reads
exactly three or four tab-separated columns, which the shipped example table (trailing tabs, CRLF) is not.  Hence the
inputs written here: ``example3.tsv`` = the (id, ts, te) columns of example_data/example_dataTAD.txt and ``trend.tsv`` =
two synthetic trend columns (a noisy ramp and a hump), one row per unit bin from the first birth year.
trend_rate.py writes next to its input (:110), so inputs are copied to a scratch directory first.
"""
from __future__ import annotations

import gzip
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = os.environ.get("LITERATE_REFERENCE", "/root/reference")
GOLD = os.path.join(REPO, "tests", "golden", "trendrate")
sys.path.insert(0, REPO)


def write_inputs():
    os.makedirs(GOLD, exist_ok=True)
    rows = [l.split() for l in open(os.path.join(REF, "example_data/example_dataTAD.txt")).read().splitlines()[1:] if l.strip()]
    with open(os.path.join(GOLD, "example3.tsv"), "w") as fh:
        fh.write("id\tts\tte\n")
        for r in rows:
            fh.write("%s\t%s\t%s\n" % (r[1], r[2], r[3]))
    rng = np.random.Generator(np.random.Philox(20260101))
    with open(os.path.join(GOLD, "trend.tsv"), "w") as fh:     # 1994 .. 2018: one row per edge interval of the example table
        fh.write("year\tramp\thump\n")
        for i in range(25):
            fh.write("%d\t%r\t%r\n" % (1994 + i, float(10 + 0.7 * i + rng.normal()), float(5 + 3 * np.exp(-((i - 9) / 5.0) ** 2) + 0.2 * rng.normal())))
    with open(os.path.join(GOLD, "trend_metal.tsv"), "w") as fh:   # 1968 .. 2000 for the metal-band table
        fh.write("year\tramp\n")
        for i in range(33):
            fh.write("%d\t%r\n" % (1968 + i, float(i * i / 40.0 + rng.normal())))


def run_reference(inputs, data, args):
    """Run trend_rate.py on scratch copies; returns ({file name: bytes}, seconds)."""
    work = tempfile.mkdtemp(prefix="lr_trend_")
    try:
        for src in inputs:
            dst = os.path.join(work, os.path.basename(src).replace(".gz", ""))
            if src.endswith(".gz"):
                with gzip.open(src, "rb") as a, open(dst, "wb") as b:
                    b.write(a.read())
            else:
                shutil.copy(src, dst)
        cmd = [sys.executable, os.path.join(REF, "trend_rate.py"), "-d", os.path.join(work, data)] + \
              [os.path.join(work, a) if a.endswith(".tsv") else a for a in args]
        t0 = time.time()
        p = subprocess.run(cmd, cwd=work, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError(p.stderr[-2000:])
        logs = {}
        for f in sorted(os.listdir(work)):
            if f.endswith(".trendrate.log"):
                with open(os.path.join(work, f), "rb") as fh:
                    logs[f] = fh.read()
        return logs, time.time() - t0
    finally:
        shutil.rmtree(work, ignore_errors=True)


EX = [os.path.join(GOLD, "example3.tsv"), os.path.join(GOLD, "trend.tsv")]
METAL = [os.path.join(REPO, "tests", "golden", "inputs", "metal_bands_1.tsv.gz"), os.path.join(GOLD, "trend_metal.tsv")]
JOBS = [
    ("ex_ramp", EX, "example3.tsv", ["-trend_data", "trend.tsv", "-trend_index", "1", "-n", "3001", "-s", "50", "-seed", "1"]),
    ("ex_hump", EX, "example3.tsv", ["-trend_data", "trend.tsv", "-trend_index", "2", "-n", "3001", "-s", "50", "-seed", "2"]),
    ("ex_constB", EX, "example3.tsv", ["-trend_data", "trend.tsv", "-trend_index", "1", "-n", "2001", "-s", "50", "-seed", "3", "-const_B", "1"]),
    ("ex_constD", EX, "example3.tsv", ["-trend_data", "trend.tsv", "-trend_index", "1", "-n", "2001", "-s", "50", "-seed", "4", "-const_D", "1"]),
    ("ex_nodeath", EX, "example3.tsv", ["-trend_data", "trend.tsv", "-trend_index", "2", "-n", "2001", "-s", "50", "-seed", "5", "-no_death", "1"]),
    ("ex_rmfirst", EX, "example3.tsv", ["-trend_data", "trend.tsv", "-trend_index", "1", "-n", "2001", "-s", "50", "-seed", "6", "-rm_first_bin", "1"]),
    ("ex_jitter0", EX, "example3.tsv", ["-trend_data", "trend.tsv", "-trend_index", "1", "-n", "1001", "-s", "50", "-seed", "7", "-death_jitter", "0"]),
    ("metal_ramp", METAL, "metal_bands_1.tsv", ["-trend_data", "trend_metal.tsv", "-trend_index", "1", "-n", "1501", "-s", "50", "-seed", "7"]),
]


def kat():
    write_inputs()
    manifest = []
    for tag, inputs, data, args in JOBS:
        if tag == "ex_jitter0":
            # with -death_jitter 0 the table ends one year earlier: the trend table needs one row fewer (trend_rate.py:82
            # broadcasts trend[:-1] against the n_bins statistics) -- write a shortened copy
            rows = open(os.path.join(GOLD, "trend.tsv")).read().splitlines()
            with open(os.path.join(GOLD, "trend_short.tsv"), "w") as fh:
                fh.write("\n".join(rows[:-1]) + "\n")
            inputs = [inputs[0], os.path.join(GOLD, "trend_short.tsv")]
            args = [a.replace("trend.tsv", "trend_short.tsv") for a in args]
        logs, dt = run_reference(inputs, data, args)
        d = os.path.join(GOLD, tag)
        os.makedirs(d, exist_ok=True)
        for f, b in logs.items():
            with open(os.path.join(d, f), "wb") as fh:
                fh.write(b)
        manifest.append({"tag": tag, "inputs": [os.path.basename(i) for i in inputs], "data": data, "args": args, "files": sorted(logs)})
        print(tag, "%.1fs" % dt, sorted(logs))
    with open(os.path.join(GOLD, "manifest.json"), "w") as fh:
        json.dump(manifest, fh, indent=1)


def _summarise(raw, burnin=0.2):
    import io
    head = raw.split(b"\n")[0].decode().rstrip("\r").split("\t")
    t = np.loadtxt(io.BytesIO(raw), skiprows=1)
    post = t[int(burnin * len(t)):]
    col = {h: i for i, h in enumerate(head)}
    out = {"n_samples": int(len(post))}
    for k in ["likelihood", "likelihood_birth", "likelihood_death", "prior", "l_min", "m_min", "alpha", "beta", "delta", "gamma"]:
        out[k + "_mean"] = float(post[:, col[k]].mean())
    nb = sum(1 for h in head if h.startswith("l_") and h[2:].isdigit())
    out["birth_rate_mean"] = post[:, col["l_0"]:col["l_0"] + nb].mean(axis=0).tolist()
    out["death_rate_mean"] = post[:, col["m_0"]:col["m_0"] + nb].mean(axis=0).tolist()
    return out


def posterior(n_chains=int(os.environ.get("LR_GOLDEN_CHAINS", "32")), n_iter=400001, s=200):
    """Posterior summaries of ``n_chains`` unmodified reference chains per configuration (distinct seeds), for the
    distributional parity of the device chains (Philox cannot replay MT19937)."""
    write_inputs()
    confs = [("ex_ramp", EX, "example3.tsv", ["-trend_data", "trend.tsv", "-trend_index", "1"]),
             ("ex_hump_constD", EX, "example3.tsv", ["-trend_data", "trend.tsv", "-trend_index", "2", "-const_D", "1"])]
    os.makedirs(os.path.join(GOLD, "posterior"), exist_ok=True)
    for tag, inputs, data, args in confs:
        def one(seed):
            logs, dt = run_reference(inputs, data, args + ["-n", str(n_iter), "-s", str(s), "-seed", str(seed)])
            (raw,) = logs.values()
            return _summarise(raw), dt
        t0 = time.time()
        with ThreadPoolExecutor(max_workers=int(os.environ.get("LR_GOLDEN_WORKERS", str(max(1, (os.cpu_count() or 2) - 1))))) as ex:
            res = list(ex.map(one, range(101, 101 + n_chains)))
        with open(os.path.join(GOLD, "posterior", tag + ".json"), "w") as fh:
            json.dump({"tag": tag, "args": args, "n_iter": n_iter, "sample_every": s, "burnin": 0.2,
                       "seeds": list(range(101, 101 + n_chains)), "chains": [r[0] for r in res]}, fh)
        print(tag, "%.0fs wall, %.0fs per chain" % (time.time() - t0, np.mean([r[1] for r in res])))


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "kat"
    {"kat": kat, "posterior": posterior}[what]()
