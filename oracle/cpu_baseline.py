"""CPU baseline legs of bench.py.  TEST/MEASUREMENT INFRASTRUCTURE ONLY -- never on the product path.

The reference (LiteRateForward.py) is a Python script that cannot be imported and does not exist on the
GPU box, so the CPU arm times the oracle port with the reference's own cost profile: the binning uses
the same NumPy calls as precompute_events/get_br (:111-123) and the chain calls scipy.stats for the
priors (``exact_scipy=True``) exactly where the reference does (:201-202, :22-23) -- the oracle chain
with these settings reproduces the reference's log files byte for byte (tests/test_oracle_golden.py).

One *sample* = what one host core does in one step of the CPU arm:
  * `bins` unit bins of one 1M-lineage replicate through the reference's per-bin formulation, and
  * `iters` RJMCMC iterations of one chain on that replicate's statistics (sampling every `s_freq`).
The whole workload (n_rep replicates x n_bins bins, n_chains x n_iter iterations) is `scale` samples, so
    value = n_chains * n_iter / (t_bin * n_bins/bins * n_rep + t_loop * n_iter/iters * n_chains) * cores
with t_bin, t_loop the slowest worker's times (workers run concurrently, one per core).
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _REPO not in sys.path:
    sys.path.insert(0, _REPO)


def _worker(job):
    from oracle import literate_oracle as O
    from literate_b200 import synth            # numpy-only generator of the synthetic tables
    n, replicate, bins, iters, s_freq, seed, model = job
    ts, te = synth.syn_int(n, replicate=replicate)
    first = int(ts.min())
    nb = int(te.max()) - first
    rng = np.random.default_rng(seed)
    which = rng.choice(nb, size=min(bins, nb), replace=False)
    t0 = time.perf_counter()
    chk = [O.events_in_bin_asref(ts, te, first + int(j), first + int(j) + 1) for j in which]
    t_bin = time.perf_counter() - t0
    st = O.bin_stats_fast(ts, te)              # untimed: only feeds the chain sample
    for j, c in zip(which, chk):
        assert (st.sp[j], st.ex[j], st.br[j]) == c
    lin = O.Lineages(ts, te, ts.min(), te.max(), 0)
    cfg = O.ChainConfig(n_iterations=iters, s_freq=s_freq, model_BDI=model, exact_scipy=True)
    t0 = time.perf_counter()
    logs = O.run_chain(lin, st, cfg, seed)
    t_loop = time.perf_counter() - t0
    return t_bin, t_loop, logs.n_lik_evals


def run_sample(cores, n_lineages, n_bins, n_rep, n_chains, n_iter, s_freq, bins=2, iters=5000, seed0=1, model=0):
    """One step of the CPU arm on `cores` worker processes (plain subprocesses: safe next to an initialised
    CUDA context); returns a dict with the extrapolated whole-workload rate."""
    import json
    import subprocess
    t0 = time.perf_counter()
    procs = []
    for r in range(cores):
        job = [n_lineages, r, bins, iters, s_freq, seed0 + r, model]
        procs.append(subprocess.Popen([sys.executable, os.path.abspath(__file__), json.dumps(job)], stdout=subprocess.PIPE,
                                      cwd=_REPO, env=dict(os.environ, OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1", MKL_NUM_THREADS="1")))
    res = []
    for p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("CPU baseline worker failed")
        res.append(json.loads(out.decode().strip().splitlines()[-1]))
    wall = time.perf_counter() - t0
    t_bin = max(r[0] for r in res)
    t_loop = max(r[1] for r in res)
    evals = sum(r[2] for r in res)
    full = (t_bin * (n_bins / bins) * n_rep + t_loop * (n_iter / iters) * n_chains) / cores
    return {
        "wall_s": wall, "t_bin_s": t_bin, "t_loop_s": t_loop,
        "it_per_s_loop_1core": iters / t_loop, "s_per_bin_1core": t_bin / bins,
        "lik_evals_per_it": evals / (cores * iters),
        "whole_workload_s": full, "value": n_chains * n_iter / full,
    }


if __name__ == "__main__":
    import json
    print(json.dumps(_worker(tuple(json.loads(sys.argv[1])))))
