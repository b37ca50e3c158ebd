"""K7 (DDRate chains) timing probe: python tools/k7_bench.py [chains,chains,...] [iters] [n_bins,...] [m_birth] [m_death]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from literate_b200 import engine as E, ddrate as DD

chains_list = [int(c) for c in (sys.argv[1] if len(sys.argv) > 1 else "256,4096").split(",")]
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
bins_list = [int(c) for c in (sys.argv[3] if len(sys.argv) > 3 else "24,200").split(",")]
mb = int(sys.argv[4]) if len(sys.argv) > 4 else 3
md = int(sys.argv[5]) if len(sys.argv) > 5 else 2
dev = E.Device(0)
tdev = torch.device("cuda:0")
rng = np.random.default_rng(1)
res = {}
for nb in bins_list:
    t = np.arange(nb)
    br = 20 + 400 / (1 + np.exp(-0.3 * (t - nb / 2))) + rng.uniform(0, 5, nb)
    sp = rng.poisson(br * 0.2); ex = rng.poisson(br * 0.1)
    gts = np.sort(rng.uniform(0, nb * .7, 40)).round(); gte = np.minimum(gts + rng.integers(1, nb, 40), nb) + .5
    for nch in chains_list:
        ch = DD.DDChains(dev, sp, ex, br, 0.0, nb + 1.5, mb, md, gts, gte, nch, 1)
        ch.run(2000)
        nrec = ch.records_per_run(iters, 1000)
        rec = torch.empty((nrec, nch, ch.rec_doubles), dtype=torch.float64, device=tdev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ch.run_device(iters, 1000, rec); b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        st = ch.state()
        res[f"k7_{nb}bins_{nch}"] = {"it_per_s": nch * iters / (ms * 1e-3), "ms": ms, "ns_per_it_per_chain": ms * 1e6 / iters,
                                     "acc_rate": float(st[:, 16].sum() / st[:, 15].sum())}
        print(nb, "bins", nch, "chains", res[f"k7_{nb}bins_{nch}"], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/k7_bench.json", "w"), indent=1)
