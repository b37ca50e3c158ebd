#!/usr/bin/env python
"""tests/golden/proportion/: a synthetic two-series table (the reference ships no input for `-proportion 1`) and the log files the
UNMODIFIED LiteRateForward-proportion.py writes for it.  TEST INFRASTRUCTURE ONLY; needs /root/reference.

    python oracle/make_golden_proportion.py kat         # seconds: exact logs of three short runs
    python oracle/make_golden_proportion.py posterior   # minutes: 32 reference chains -> posterior summaries
"""
import json
import os
import shutil
import subprocess
import sys
import tempfile
from concurrent.futures import ThreadPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = os.environ.get("LITERATE_REFERENCE", "/root/reference")
GOLD = os.path.join(REPO, "tests", "golden", "proportion")
sys.path.insert(0, REPO)


def make_table(path):
    """Two event series over 1975..2019: the first one dense with a changing rate, the second one sparser, with gap years
    (interpolated by the script) and empty cells."""
    rng = np.random.default_rng(20260317)
    years = np.arange(1975, 2020)
    lam1 = np.where(years < 1990, 6.0, np.where(years < 2005, 14.0, 9.0))
    lam2 = np.where(years < 1998, 3.0, 7.0)
    s1 = np.repeat(years, rng.poisson(lam1))
    c2 = rng.poisson(lam2); c2[[3, 4, 17, 30]] = 0; c2[0] = max(c2[0], 1)
    s2 = np.repeat(years, c2)
    rng.shuffle(s1); rng.shuffle(s2)
    n = max(len(s1), len(s2))
    with open(path, "w") as fh:
        fh.write("id\tnumerator_year\tdenominator_year\n")
        for i in range(n):
            fh.write("%d\t%s\t%s\n" % (i, s1[i] if i < len(s1) else "", s2[i] if i < len(s2) else ""))


def run_reference(src, args):
    work = tempfile.mkdtemp(prefix="lr_prop_")
    try:
        name = os.path.basename(src)
        shutil.copy(src, os.path.join(work, name))
        p = subprocess.run([sys.executable, os.path.join(REF, "LiteRateForward-proportion.py"), "-d", os.path.join(work, name), "-proportion", "1"] + args,
                           cwd=work, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError(p.stderr[-3000:])
        d = os.path.join(work, "literate_mcmc_logs")
        return {f: open(os.path.join(d, f), "rb").read() for f in sorted(os.listdir(d))}
    finally:
        shutil.rmtree(work, ignore_errors=True)


def kat():
    os.makedirs(GOLD, exist_ok=True)
    src = os.path.join(GOLD, "two_series.tsv")
    make_table(src)
    manifest = []
    for tag, args in (("pr_default", ["-n", "3000", "-s", "25", "-p", "1000000", "-seed", "1"]),
                      ("pr_noadq_constdeath", ["-n", "3000", "-s", "25", "-p", "1000000", "-seed", "4", "-calc_adequacy", "0", "-const_death_rate", "1"]),
                      ("pr_fixedpoi", ["-n", "2000", "-s", "25", "-p", "1000000", "-seed", "9", "-Poisson_prior", "2", "-use_rate_HP", "0", "-death_jitter", "0.25"])):
        logs = run_reference(src, args)
        d = os.path.join(GOLD, tag)
        os.makedirs(d, exist_ok=True)
        for f, b in logs.items():
            with open(os.path.join(d, f), "wb") as fh:
                fh.write(b)
        manifest.append({"tag": tag, "args": args, "files": sorted(logs)})
        print(tag, sorted(logs))
    json.dump(manifest, open(os.path.join(GOLD, "manifest.json"), "w"), indent=1)


def posterior(n_chains=int(os.environ.get("LR_GOLDEN_CHAINS", "32")), n_it=200001, s=100):
    from oracle import literate_oracle as O
    from oracle import proportion_oracle as PO
    import io
    src = os.path.join(GOLD, "two_series.tsv")
    lin = PO.read_series(src)

    def one(seed):
        logs = run_reference(src, ["-n", str(n_it), "-s", str(s), "-p", "100000000", "-seed", str(seed)])
        stem = "two_series_PR_seed%d" % seed
        mc = np.loadtxt(io.BytesIO(logs[stem + "_mcmc.log"]), skiprows=1)
        b = int(0.2 * len(mc)); post = mc[b:]
        res = {"seed": seed, "n_iterations": n_it, "s_freq": s, "n_samples": int(len(post)),
               "K_l": O.k_pmf(mc[:, 6]), "K_m": O.k_pmf(mc[:, 7]), "lik_mean": float(post[:, 2].mean()), "lik_var": float(post[:, 2].var()),
               "prior_mean": float(post[:, 3].mean()), "lambda_avg": float(post[:, 4].mean()), "mu_avg": float(post[:, 5].mean()),
               "gamma_hp_l": float(post[:, 10].mean()), "gamma_hp_m": float(post[:, 11].mean()), "poisson_hp": float(post[:, 12].mean())}
        for tag, key in (("sp_rates", "birth"), ("ex_rates", "death")):
            rows = [np.array(l.split(), dtype=np.float64) for l in logs[stem + "_" + tag + ".log"].decode().split("\n") if l.strip()]
            res[key + "_rate_mean"] = O.marginal_rates(rows, lin.end_time, lin.start_time, 0.2).mean(axis=0).tolist()
        return res
    with ThreadPoolExecutor(int(os.environ.get("LR_GOLDEN_WORKERS", "6"))) as ex:
        chains = list(ex.map(one, [801 + i for i in range(n_chains)]))
    json.dump({"data": "tests/golden/proportion/two_series.tsv", "generator": "unmodified LiteRateForward-proportion.py -proportion 1", "burnin": 0.2,
               "chains": chains}, open(os.path.join(GOLD, "posterior.json"), "w"), indent=1)
    print("posterior done")


if __name__ == "__main__":
    {"kat": kat, "posterior": posterior}[sys.argv[1] if len(sys.argv) > 1 else "kat"]()
