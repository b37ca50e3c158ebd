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


# DDRatev3.py:48 reads three names that only the -m_birth 3 branch defines (:35-37), so the shipped script stops with a NameError
# for every other -m_birth before it samples anything.  The jobs marked shim=True execute its UNMODIFIED text (runpy) with those
# three names pre-bound in builtins -- line 48 then prints the empirical rates of a dummy table and nothing else ever reads
# them -- which is what pins the oracle's -m_birth 0 / 1 / 2 branches (:86-99) to the reference's own arithmetic.


def run_reference(inputs, data, args, shim=False):
    work = tempfile.mkdtemp(prefix="lr_dd_")
    try:
        for src in inputs:
            dst = os.path.join(work, os.path.basename(src).replace(".gz", ""))
            if src.endswith(".gz"):
                with gzip.open(src, "rb") as a, open(dst, "wb") as b:
                    b.write(a.read())
            else:
                shutil.copy(src, dst)
        script = os.path.join(REF, "DDRatev3.py")
        if shim:
            with open(os.path.join(work, "_shim.py"), "w") as fh:
                fh.write("import builtins, sys, runpy, numpy as np\n"
                         "builtins.GN_SPEC = np.array([1.0]); builtins.GN_EXTI = np.array([1.0]); builtins.GDT = np.array([1.0])\n"
                         "script = sys.argv[1]; sys.argv = [script] + sys.argv[2:]\n"
                         "sys.path.insert(0, __import__('os').path.dirname(script))\n"
                         "runpy.run_path(script, run_name='__main__')\n")
            head = [sys.executable, os.path.join(work, "_shim.py"), script]
        else:
            head = [sys.executable, script]
        cmd = head + ["-d", os.path.join(work, data)] + [os.path.join(work, a) if a.endswith(".tsv") else a for a in args]
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
SHIM_JOBS = [      # -m_birth 0 / 1 / 2: run through the three-name shim described above run_reference
    ("ex_ll_mddn", [EX[0]], "example3.tsv", ["-m_birth", "0", "-n", "3001", "-s", "50", "-seed", "11"]),
    ("ex_ldd_mdd", [EX[0]], "example3.tsv", ["-m_birth", "1", "-m_death", "1", "-n", "3001", "-s", "50", "-seed", "12"]),
    ("ex_lddn_ml", [EX[0]], "example3.tsv", ["-m_birth", "2", "-m_death", "0", "-n", "3001", "-s", "50", "-seed", "13"]),
    ("ex_ldd_mddn", [EX[0]], "example3.tsv", ["-m_birth", "1", "-m_death", "2", "-n", "2001", "-s", "50", "-seed", "14"]),
]
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
    for tag, inputs, data, args, shim in [j + (False,) for j in JOBS] + [j + (True,) for j in SHIM_JOBS]:
        logs, dt = run_reference(inputs, data, args, shim=shim)
        d = os.path.join(GOLD, tag)
        os.makedirs(d, exist_ok=True)
        for f, b in logs.items():
            with open(os.path.join(d, f), "wb") as fh:
                fh.write(b)
        manifest.append({"tag": tag, "inputs": [os.path.basename(i) for i in inputs], "data": data, "args": args, "files": sorted(logs),
                         "shim": shim})
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
