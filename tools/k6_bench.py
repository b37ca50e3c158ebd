"""K6 (TrendRate chains) timing probe: python tools/k6_bench.py [chains,chains,...] [iters] [n_bins,n_bins,...]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from literate_b200 import engine as E, trend as TR

chains_list = [int(c) for c in (sys.argv[1] if len(sys.argv) > 1 else "256,4096").split(",")]
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
bins_list = [int(c) for c in (sys.argv[3] if len(sys.argv) > 3 else "24,200").split(",")]
dev = E.Device(0)
tdev = torch.device("cuda:0")
rng = np.random.default_rng(1)
res = {}
for nb in bins_list:
    br = rng.uniform(50, 500, nb)
    trend = np.clip(np.linspace(0, 1, nb) + 0.05 * rng.normal(size=nb), 1e-15, 1.0)
    sp = rng.poisson(br * (0.1 + 0.2 * trend)); ex = rng.poisson(br * (0.05 + 0.1 * trend))
    for nch in chains_list:
        ch = TR.TrendChains(dev, sp, ex, br, trend, nch, 1)
        ch.run(2000)
        nrec = ch.records_per_run(iters, 1000)
        rec = torch.empty((nrec, nch, ch.rec_doubles), dtype=torch.float64, device=tdev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ch.run_device(iters, 1000, rec); b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        st = ch.state()
        res[f"k6_{nb}bins_{nch}"] = {"it_per_s": nch * iters / (ms * 1e-3), "ms": ms, "ns_per_it_per_chain": ms * 1e6 / iters,
                                     "acc_rate": float(st[:, 10].sum() / st[:, 9].sum())}
        print(nb, "bins", nch, "chains", res[f"k6_{nb}bins_{nch}"], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/k6_bench.json", "w"), indent=1)
