"""K3 timing probe (development aid): python tools/k3_bench.py [chains,chains,...] [iters] [sample_every] [loop variants, e.g. 1,2]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from literate_b200 import engine as E, synth

chains_list = [int(c) for c in (sys.argv[1] if len(sys.argv) > 1 else "256,4096").split(",")]
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
s_every = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
variants = [int(v) for v in (sys.argv[4] if len(sys.argv) > 4 else "0").split(",")]
dev = E.Device(0)
tdev = torch.device("cuda:0")
n, nb, n_rep = 1_000_000, 200, 4
ts, te = synth.syn_int_device(n, n_rep, tdev)
sp, ex, br = dev.bin_stats_device(ts[:, :n], te[:, :n], 1800, nb)
torch.cuda.synchronize()
ds = E.Dataset.from_device(dev, sp, ex, br, 0, 1800.0, 2000.5)
res = {}
for nch, variant in [(n_, v_) for n_ in chains_list for v_ in variants]:
    ch = E.Chains(ds, nch, 1, cfg=E.default_config(0, loop_variant=variant), rep_of_chain=np.arange(nch) % n_rep)
    ch.run(2000)                                   # burn-in: K grows to its stationary range
    nrec = ch.records_per_run(iters, s_every)
    rec = torch.empty((nrec, nch, E.LR_REC_DOUBLES), dtype=torch.float64, device=tdev)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); ch.run_device(iters, s_every, rec); b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    cnt = ch.counters().sum(0)
    st = ch.state()
    res[f"k3_{nch}_v{variant}"] = {"it_per_s": nch * iters / (ms * 1e-3), "ms": ms, "ns_per_it_per_chain": ms * 1e6 / iters,
                        "acc_rate": float(cnt[1] / cnt[0]), "K_l": float(st[:, E.REC_KL].mean()), "K_m": float(st[:, E.REC_KM].mean())}
    print(nch, "variant", variant, res[f"k3_{nch}_v{variant}"], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/k3_bench.json", "w"), indent=1)
