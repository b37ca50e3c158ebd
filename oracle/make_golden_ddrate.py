#!/usr/bin/env python
"""Generate tests/golden/ddrate/ by running the UNMODIFIED DDRatev3.py in this container.

TEST INFRASTRUCTURE ONLY.  Needs /root/reference (read-only, absent on the GPU box); everything it produces is committed.

    python oracle/make_golden_ddrate.py kat         # seconds: exact logs of short reference runs
    python oracle/make_golden_ddrate.py posterior   # minutes: 8 reference chains per configuration -> posterior summaries

As shipped DDRatev3.py only runs with ``-m_birth 3 -g <genre table>`` (NameError at :48 otherwise, see
oracle/ddrate_oracle.py); the tree holds no genre table, so a small synthetic one (``genres.tsv``: id, first year, last
year of 14 "genres" inside the window of the example table) is written here next to the three-column copy of the example
table that oracle/make_golden_trend.py makes.  The script writes next to its input (:156, :170): inputs are copied to a
scratch directory first.
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
GOLD = os.path.join(REPO, "tests", "golden", "ddrate")
TREND = os.path.join(REPO, "tests", "golden", "trendrate")
sys.path.insert(0, REPO)


def write_inputs():
    os.makedirs(GOLD, exist_ok=True)
    rng = np.random.Generator(np.random.Philox(20260102))
    with open(os.path.join(GOLD, "genres.tsv"), "w") as fh:
        fh.write("id\tts\tte\n")
        for i in range(14):
            ts = 1994 + int(rng.integers(0, 16))
            te = min(2017, ts + 1 + int(rng.integers(0, 20)))
            fh.write("%d\t%d\t%d\n" % (i, ts, te))
    with open(os.path.join(GOLD, "genres_metal.tsv"), "w") as fh:
        fh.write("id\tts\tte\n")
        for i in range(40):
            ts = 1968 + int(rng.integers(0, 25))
            te = min(2000, ts + 1 + int(rng.integers(0, 25)))
            fh.write("%d\t%d\t%d\n" % (i, ts, te))


def run_reference(inputs, data, args):
    work = tempfile.mkdtemp(prefix="lr_dd_")
    try:
        for src in inputs:
            dst = os.path.join(work, os.path.basename(src).replace(".gz", ""))
            if src.endswith(".gz"):
                with gzip.open(src, "rb") as a, open(dst, "wb") as b:
                    b.write(a.read())
            else:
                shutil.copy(src, dst)
        cmd = [sys.executable, os.path.join(REF, "DDRatev3.py"), "-d", os.path.join(work, data)] + \
              [os.path.join(work, a) if a.endswith(".tsv") else a for a in args]
        t0 = time.time()
        p = subprocess.run(cmd, cwd=work, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError(p.stderr[-2000:])
        logs = {}
        for f in sorted(os.listdir(work)):
            if f.endswith(".log"):
                with open(os.path.join(work, f), "rb") as fh:
                    logs[f] = fh.read()
        return logs, time.time() - t0
    finally:
        shutil.rmtree(work, ignore_errors=True)


EX = [os.path.join(TREND, "example3.tsv"), os.path.join(GOLD, "genres.tsv")]
METAL = [os.path.join(REPO, "tests", "golden", "inputs", "metal_bands_1.tsv.gz"), os.path.join(GOLD, "genres_metal.tsv")]
G = ["-m_birth", "3", "-g", "genres.tsv"]
JOBS = [
    ("ex_g_mddn", EX, "example3.tsv", G + ["-n", "3001", "-s", "50", "-seed", "1"]),
    ("ex_g_mdd", EX, "example3.tsv", G + ["-m_death", "1", "-n", "3001", "-s", "50", "-seed", "2"]),
    ("ex_g_ml", EX, "example3.tsv", G + ["-m_death", "0", "-n", "3001", "-s", "50", "-seed", "3"]),
    ("ex_g_rmfirst", EX, "example3.tsv", G + ["-n", "2001", "-s", "50", "-seed", "4", "-rm_first_bin", "1"]),
    ("metal_g_mddn", METAL, "metal_bands_1.tsv", ["-m_birth", "3", "-g", "genres_metal.tsv", "-n", "1501", "-s", "50", "-seed", "7"]),
]


def kat():
    write_inputs()
    manifest = []
    for tag, inputs, data, args in JOBS:
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


def _summarise(raw, n_bins, burnin=0.2):
    """Column positions, not names: the reference's header is one name short of its rows (:159-161 vs :285)."""
    import io
    t = np.loadtxt(io.BytesIO(raw), skiprows=1)
    post = t[int(burnin * len(t)):]
    names = ["likelihood", "likelihood_birth", "likelihood_death", "prior", "l_f", "l_mul", "k", "x0_abs", "div_0", "K_max",
             "m_mul", "nuB", "nuD", "g_l1", "g_l2", "likelihood_genre"]
    out = {"n_samples": int(len(post))}
    for i, k in enumerate(names):
        out[k + "_mean"] = float(post[:, 2 + i].mean())
    c = 2 + len(names)
    for i, k in enumerate(["birth_rate", "death_rate", "niche", "niche_frac"]):
        out[k + "_mean"] = post[:, c + i * n_bins:c + (i + 1) * n_bins].mean(axis=0).tolist()
    return out


def posterior(n_chains=int(os.environ.get("LR_GOLDEN_CHAINS", "32")), n_iter=300001, s=200):
    write_inputs()
    confs = [("ex_g_mddn", EX, "example3.tsv", G, 24), ("ex_g_mdd", EX, "example3.tsv", G + ["-m_death", "1"], 24)]
    os.makedirs(os.path.join(GOLD, "posterior"), exist_ok=True)
    for tag, inputs, data, args, nb in confs:
        def one(seed):
            logs, dt = run_reference(inputs, data, args + ["-n", str(n_iter), "-s", str(s), "-seed", str(seed)])
            raw = [b for f, b in logs.items() if not f.endswith(".div.log")][0]
            return _summarise(raw, nb), dt
        t0 = time.time()
        with ThreadPoolExecutor(max_workers=int(os.environ.get("LR_GOLDEN_WORKERS", str(max(1, (os.cpu_count() or 2) - 1))))) as ex:
            res = list(ex.map(one, range(201, 201 + n_chains)))
        with open(os.path.join(GOLD, "posterior", tag + ".json"), "w") as fh:
            json.dump({"tag": tag, "args": args, "n_iter": n_iter, "sample_every": s, "burnin": 0.2,
                       "seeds": list(range(201, 201 + n_chains)), "chains": [r[0] for r in res]}, fh)
        print(tag, "%.0fs wall, %.0fs per chain" % (time.time() - t0, np.mean([r[1] for r in res])))


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "kat"
    {"kat": kat, "posterior": posterior}[what]()
