#!/usr/bin/env python
"""Generate tests/golden/ by running the UNMODIFIED reference in this container.

TEST INFRASTRUCTURE ONLY.  Needs /root/reference (read-only, not present on the GPU box);
everything it produces is committed so that no test reads /root/reference at run time.

    python oracle/make_golden.py kat         # seconds: exact log files of short reference runs
    python oracle/make_golden.py posterior   # minutes: 8 reference chains per data set -> summaries
    python oracle/make_golden.py posterior_flags   # ~10 min: 8 reference chains per non-default flag set (example TAD)
    python oracle/make_golden.py prior_only   # ~2 min: 8 ORACLE chains on the example window with all statistics zero
    python oracle/make_golden.py plots       # ~30 s: the reference's plotRJforward.v3.py on two of the kat runs -> .r files
    python oracle/make_golden.py averager    # seconds: utilities/imputation_averager.py on five synthetic imputation div.log files
    python oracle/make_golden.py posterior_syn   # ~10 min: 32 ORACLE chains on the bench statistics (syn-int 1M lineages x 200 bins)

posterior / posterior_flags take the number of reference chains per fixture from LR_GOLDEN_CHAINS (default 32) and run at
most LR_GOLDEN_WORKERS (default: cores - 1) reference processes at a time.

The reference writes next to its input (LiteRateForward.py:479-491), so inputs are copied to
a scratch directory first.
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

N_CHAINS = int(os.environ.get("LR_GOLDEN_CHAINS", "32"))
N_WORKERS = int(os.environ.get("LR_GOLDEN_WORKERS", str(max(1, (os.cpu_count() or 2) - 1))))

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = os.environ.get("LITERATE_REFERENCE", "/root/reference")
GOLD = os.path.join(REPO, "tests", "golden")
sys.path.insert(0, REPO)

INPUTS = {
    "example_dataTAD.txt": "example_data/example_dataTAD.txt",
    "example_dataTBP.txt": "example_data/example_dataTBP.txt",
    "metal_bands_1.tsv": "example_data/metal_bands/single_run/metal_bands_1.tsv",
}
SUFFIX = {0: "_BD", 1: "_ID", 2: "_BDk", 3: "_BDd"}


def run_reference(src, args, keep_dir=None):
    """Run LiteRateForward.py on a scratch copy of `src`; return {tag: bytes} of its logs."""
    work = tempfile.mkdtemp(prefix="lr_ref_")
    try:
        name = os.path.basename(src)
        shutil.copy(src, os.path.join(work, name))
        cmd = [sys.executable, os.path.join(REF, "LiteRateForward.py"), "-d", os.path.join(work, name)] + args
        t0 = time.time()
        p = subprocess.run(cmd, cwd=work, capture_output=True, text=True)
        dt = time.time() - t0
        if p.returncode != 0:
            raise RuntimeError(p.stderr[-2000:])
        logs = {}
        d = os.path.join(work, "literate_mcmc_logs")
        for f in sorted(os.listdir(d)):
            with open(os.path.join(d, f), "rb") as fh:
                logs[f] = fh.read()
        return logs, dt, p.stdout
    finally:
        shutil.rmtree(work, ignore_errors=True)


def stage_inputs():
    d = os.path.join(GOLD, "inputs")
    os.makedirs(d, exist_ok=True)
    for name, rel in INPUTS.items():
        with open(os.path.join(REF, rel), "rb") as fh:
            raw = fh.read()
        if len(raw) > 100_000:
            with gzip.GzipFile(os.path.join(d, name + ".gz"), "wb", mtime=0) as gz:
                gz.write(raw)
        else:
            with open(os.path.join(d, name), "wb") as fh:
                fh.write(raw)


def kat():
    stage_inputs()
    out = os.path.join(GOLD, "reference_logs")
    os.makedirs(out, exist_ok=True)
    manifest = []
    tad = os.path.join(REF, INPUTS["example_dataTAD.txt"])
    tbp = os.path.join(REF, INPUTS["example_dataTBP.txt"])
    metal = os.path.join(REF, INPUTS["metal_bands_1.tsv"])
    jobs = []
    for m in range(4):
        jobs.append(("tad_m%d" % m, tad, ["-n", "3000", "-s", "25", "-p", "1000000", "-seed", "1", "-model_BDI", str(m)]))
    jobs.append(("tbp_m0", tbp, ["-n", "3000", "-s", "25", "-p", "1000000", "-seed", "1", "-TBP"]))
    jobs.append(("tbp_pyrate", tbp, ["-n", "2000", "-s", "25", "-p", "1000000", "-seed", "3", "-TBP", "-pyrate_output"]))
    jobs.append(("tad_constdeath", tad, ["-n", "3000", "-s", "25", "-p", "1000000", "-seed", "5", "-const_death_rate", "1"]))
    jobs.append(("tad_constrates", tad, ["-n", "3000", "-s", "25", "-p", "1000000", "-seed", "6", "-const_rates", "1", "-calc_adequacy", "0"]))
    jobs.append(("tad_fixedpoi_nohp", tad, ["-n", "3000", "-s", "25", "-p", "1000000", "-seed", "7", "-Poisson_prior", "2", "-use_rate_HP", "0",
                                            "-update_fraction", "0.5", "-out", "_x"]))
    jobs.append(("tad_jitter0", tad, ["-n", "1000", "-s", "25", "-p", "1000000", "-seed", "8", "-death_jitter", "0", "-model_BDI", "3"]))
    jobs.append(("metal_m0", metal, ["-n", "1500", "-s", "25", "-p", "1000000", "-seed", "7"]))
    jobs.append(("metal_m3", metal, ["-n", "300", "-s", "25", "-p", "1000000", "-seed", "7", "-model_BDI", "3"]))
    for tag, src, args in jobs:
        logs, dt, _ = run_reference(src, args)
        d = os.path.join(out, tag)
        os.makedirs(d, exist_ok=True)
        for f, b in logs.items():
            with open(os.path.join(d, f), "wb") as fh:
                fh.write(b)
        manifest.append({"tag": tag, "input": os.path.basename(src), "args": args, "files": sorted(logs)})
        print(tag, "%.1fs" % dt, sorted(logs))
    with open(os.path.join(out, "manifest.json"), "w") as fh:
        json.dump(manifest, fh, indent=1)


def _summarise(logs, stem, burnin=0.2):
    """Per-chain posterior summaries from the reference's own log files."""
    from oracle import literate_oracle as O
    import io
    mc = np.loadtxt(io.BytesIO(logs[stem + "_mcmc.log"]), skiprows=1)
    head = logs[stem + "_mcmc.log"].split(b"\n")[0].decode().split("\t")
    n = len(mc)
    b = int(burnin * n)
    post = mc[b:]
    col = {h: i for i, h in enumerate(head)}
    start, end = post[0, col["root_age"]], post[0, col["death_age"]]
    res = {"n_samples": int(len(post)),
           "K_l": O.k_pmf(mc[:, col["K_l"]], burnin), "K_m": O.k_pmf(mc[:, col["K_m"]], burnin),
           "lik_mean": float(post[:, col["likelihood"]].mean()), "lik_var": float(post[:, col["likelihood"]].var()),
           "prior_mean": float(post[:, col["prior"]].mean()),
           "lambda_avg": float(post[:, col["lambda_avg"]].mean()), "mu_avg": float(post[:, col["mu_avg"]].mean()),
           "gamma_hp_l": float(post[:, col["gamma_rate_hp_BI"]].mean()), "gamma_hp_m": float(post[:, col["gamma_rate_hp_D"]].mean()),
           "poisson_hp": float(post[:, col["poisson_rate_hp"]].mean())}
    for tag, key in (("sp_rates", "birth"), ("ex_rates", "death")):
        rows = [np.array(l.split(), dtype=np.float64) for l in logs[stem + "_" + tag + ".log"].decode().split("\n") if l.strip()]
        mr = O.marginal_rates(rows, end, start, burnin)
        res[key + "_rate_mean"] = mr.mean(axis=0).tolist()
    return res


def plots():
    """tests/golden/plots/<tag>_RTT_plots.r: what the UNMODIFIED plotRJforward.v3.py writes for the log files of two kat runs
    (posterior-summary fixtures: K pmf, marginal rates + HPD, shift frequencies, net rate; the bf2/bf6 lines come from an
    unseeded prior simulation and differ from run to run)."""
    out = os.path.join(GOLD, "plots")
    os.makedirs(out, exist_ok=True)
    for tag in ("tad_m0", "metal_m0"):
        work = tempfile.mkdtemp(prefix="lr_plot_")
        try:
            src = os.path.join(GOLD, "reference_logs", tag)
            for f in os.listdir(src):
                shutil.copy(os.path.join(src, f), work)
            p = subprocess.run([sys.executable, os.path.join(REF, "plotRJforward.v3.py"), work], cwd=work, capture_output=True, text=True)
            r = [f for f in os.listdir(work) if f.endswith("_RTT_plots.r")]
            if not r:
                raise RuntimeError(p.stderr[-2000:])
            shutil.copy(os.path.join(work, r[0]), os.path.join(out, tag + "_RTT_plots.r"))
        finally:
            shutil.rmtree(work, ignore_errors=True)


FLAG_SETS = [   # tag, extra arguments, model suffix
    ("tad_constdeath", ["-const_death_rate", "1"], "_BD"),
    ("tad_constrates", ["-const_rates", "1"], "_BD"),
    ("tad_fixedpoi_nohp", ["-Poisson_prior", "2", "-use_rate_HP", "0"], "_BD"),
    ("tad_keiding", ["-model_BDI", "2"], "_BDk"),
    ("tad_immigration", ["-model_BDI", "1"], "_ID"),
    ("tad_keiding_dead", ["-model_BDI", "3"], "_BDd"),
]


def posterior_flags(n_chains=N_CHAINS, n_it=200001, s=100):
    """Posterior summaries of unmodified reference chains for every non-default sampler configuration (example TAD)."""
    out = os.path.join(GOLD, "posterior")
    os.makedirs(out, exist_ok=True)
    src = os.path.join(REF, INPUTS["example_dataTAD.txt"])
    for tag, extra, suffix in FLAG_SETS:
        def one(seed):
            logs, dt, _ = run_reference(src, ["-n", str(n_it), "-s", str(s), "-p", "100000000", "-seed", str(seed)] + extra)
            r = _summarise(logs, "example_dataTAD" + suffix)
            r["seed"], r["wall_s"], r["n_iterations"], r["s_freq"] = seed, dt, n_it, s
            return r
        with ThreadPoolExecutor(N_WORKERS) as ex:
            chains = list(ex.map(one, [301 + i for i in range(n_chains)]))
        with open(os.path.join(out, tag + ".json"), "w") as fh:
            json.dump({"data": "example_tad", "generator": "unmodified LiteRateForward.py " + " ".join(extra), "args": extra, "burnin": 0.2,
                       "chains": chains}, fh, indent=1)
        print(tag, "done:", [round(c["wall_s"]) for c in chains], flush=True)


def _prior_chain(seed):
    """One oracle chain (the restatement that reproduces the reference's logs byte for byte on real data) on the time window
    of the example table with EVERY sufficient statistic set to zero: the likelihood is identically 0, so the chain samples
    the prior the reference's proposals and acceptance rule define (number of shifts, shift times, rates, hyper-priors)."""
    from oracle import literate_oracle as O
    lin = O.read_lineages(os.path.join(GOLD, "inputs", "example_dataTAD.txt"))
    nb = int(lin.end_time) - int(lin.start_time)
    st = O.BinStats(int(lin.start_time), np.zeros(nb, dtype=np.int64), np.zeros(nb, dtype=np.int64), np.zeros(nb))
    cfg = O.ChainConfig(n_iterations=300001, s_freq=100, calc_adequacy=0)
    logs = O.run_chain(lin, st, cfg, seed)
    mc = np.array([[float(x) for x in l.split("\t")] for l in logs.mcmc.getvalue().splitlines()])
    b = int(0.2 * len(mc))
    post = mc[b:]
    return {"seed": seed, "K_l": O.k_pmf(mc[:, 6]), "K_m": O.k_pmf(mc[:, 7]), "lambda_avg": float(post[:, 4].mean()),
            "mu_avg": float(post[:, 5].mean()), "prior_mean": float(post[:, 3].mean()), "gamma_hp_l": float(post[:, 10].mean()),
            "poisson_hp": float(post[:, 12].mean()), "n_iterations": 300001, "s_freq": 100}


def prior_only(n_chains=8):
    from concurrent.futures import ProcessPoolExecutor
    with ProcessPoolExecutor(n_chains) as ex:
        chains = list(ex.map(_prior_chain, [501 + i for i in range(n_chains)]))
    with open(os.path.join(GOLD, "posterior", "prior_only.json"), "w") as fh:
        json.dump({"data": "example_tad window, all statistics zero", "generator": "oracle.run_chain (pinned to the reference)", "burnin": 0.2,
                   "chains": chains}, fh, indent=1)
    print("prior_only done")


def averager():
    """tests/golden/averager/: five div.log files (the reference's own writer format, LiteRateForward.py:558-564, produced by
    the oracle from synth.syn_int imputation replicates, one of them with an empty first bin) and the text the UNMODIFIED
    utilities/imputation_averager.py prints for that directory."""
    from oracle import literate_oracle as O
    from literate_b200 import synth
    out = os.path.join(GOLD, "averager")
    os.makedirs(out, exist_ok=True)
    for r in range(5):
        ts, te = synth.syn_int(3000, replicate=90 + r)
        if r == 3:
            ts[ts == ts.min()] += 1.0; ts[0] = 1800.0; te[0] = 1800.0 + 0.5      # bin 0: one birth, half a year at risk
        st = O.bin_stats_fast(ts, te)
        with open(os.path.join(out, "imputation_%d_div.log" % r), "w", newline="") as fh:
            O.write_div_log(fh, st)
    p = subprocess.run([sys.executable, os.path.join(REF, "utilities", "imputation_averager.py"), out], capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError(p.stderr[-2000:])
    with open(os.path.join(out, "expected_stdout.txt"), "w") as fh:
        fh.write(p.stdout)
    with open(os.path.join(out, "file_order.json"), "w") as fh:
        import glob as _g
        json.dump([os.path.basename(f) for f in _g.glob(out + "/*div.log")], fh)
    print(p.stdout[:400])


SYN_N, SYN_ITERS, SYN_S, SYN_BURNIN = 1_000_000, 400001, 200, 0.5


def _syn_chain(seed):
    """One oracle chain on the statistics of the bench workload: synth.syn_int(1M lineages, replicate 0) -> 200 unit bins
    (K_l ~ 11, K_m = 1).  The unmodified script would spend 3 minutes binning 1M lineages per chain before its first
    iteration; the oracle (pinned to it byte for byte on 12 flag sets) takes the statistics from bin_stats_fast, which the CPU
    tests hold to the per-bin formulation."""
    from oracle import literate_oracle as O
    from literate_b200 import synth
    ts, te = synth.syn_int(SYN_N, 0)
    st = O.bin_stats_fast(ts, te)
    lin = O.Lineages(ts=ts, te=te, start_time=float(ts.min()), end_time=float(te.max()), true_root_age=float(ts.min()))
    cfg = O.ChainConfig(n_iterations=SYN_ITERS, s_freq=SYN_S, calc_adequacy=0)
    t0 = time.time()
    logs = O.run_chain(lin, st, cfg, seed)
    mc = np.array([[float(x) for x in l.split("\t")] for l in logs.mcmc.getvalue().splitlines()])
    b = int(SYN_BURNIN * len(mc))
    post = mc[b:]
    res = {"seed": seed, "n_iterations": SYN_ITERS, "s_freq": SYN_S, "wall_s": time.time() - t0, "n_samples": int(len(post)),
           "K_l": O.k_pmf(mc[:, 6], SYN_BURNIN), "K_m": O.k_pmf(mc[:, 7], SYN_BURNIN),
           "lik_mean": float(post[:, 2].mean()), "lik_var": float(post[:, 2].var()), "prior_mean": float(post[:, 3].mean()),
           "lambda_avg": float(post[:, 4].mean()), "mu_avg": float(post[:, 5].mean()),
           "gamma_hp_l": float(post[:, 10].mean()), "gamma_hp_m": float(post[:, 11].mean()), "poisson_hp": float(post[:, 12].mean())}
    for text, key in ((logs.sp.getvalue(), "birth"), (logs.ex.getvalue(), "death")):
        rows = [np.array(l.split(), dtype=np.float64) for l in text.split("\n") if l.strip()]
        res[key + "_rate_mean"] = O.marginal_rates(rows, lin.end_time, lin.start_time, SYN_BURNIN).mean(axis=0).tolist()
    return res


def posterior_syn(n_chains=N_CHAINS):
    from concurrent.futures import ProcessPoolExecutor
    with ProcessPoolExecutor(N_WORKERS) as ex:
        chains = list(ex.map(_syn_chain, [701 + i for i in range(n_chains)]))
    with open(os.path.join(GOLD, "posterior", "syn_int_200.json"), "w") as fh:
        json.dump({"data": "synth.syn_int(1000000, replicate 0): 200 unit bins 1800..1999", "burnin": SYN_BURNIN,
                   "generator": "oracle.run_chain (pinned to the reference byte for byte) on oracle.bin_stats_fast statistics",
                   "chains": chains}, fh, indent=1)
    print("posterior_syn done:", [round(c["wall_s"]) for c in chains])


def posterior(n_chains=N_CHAINS):
    out = os.path.join(GOLD, "posterior")
    os.makedirs(out, exist_ok=True)
    sets = [
        ("example_tad", os.path.join(REF, INPUTS["example_dataTAD.txt"]), "example_dataTAD_BD", 300001, 100),
        ("metal_bands", os.path.join(REF, INPUTS["metal_bands_1.tsv"]), "metal_bands_1_BD", 400001, 200),
    ]
    for tag, src, stem, n_it, s in sets:
        def one(seed):
            logs, dt, _ = run_reference(src, ["-n", str(n_it), "-s", str(s), "-p", "100000000", "-seed", str(seed)])
            r = _summarise(logs, stem)
            r["seed"], r["wall_s"], r["n_iterations"], r["s_freq"] = seed, dt, n_it, s
            return r
        with ThreadPoolExecutor(N_WORKERS) as ex:
            chains = list(ex.map(one, [101 + i for i in range(n_chains)]))
        with open(os.path.join(out, tag + ".json"), "w") as fh:
            json.dump({"data": tag, "generator": "unmodified LiteRateForward.py, default flags", "burnin": 0.2,
                       "chains": chains}, fh, indent=1)
        print(tag, "done:", [round(c["wall_s"]) for c in chains])


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "kat"
    {"kat": kat, "posterior": posterior, "posterior_flags": posterior_flags, "prior_only": prior_only, "plots": plots,
     "posterior_syn": posterior_syn, "averager": averager}[what]()
