"""Where do the iterations of the automatic K3 policy go (teams / continuation pass)?  python tools/k3_policy_probe.py [chains]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from literate_b200 import engine as E, synth
nch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = E.Device(0); tdev = torch.device("cuda:0")
n, nb, n_rep = 1_000_000, 200, 4
ts, te = synth.syn_int_device(n, n_rep, tdev)
sp, ex, br = dev.bin_stats_device(ts[:, :n], te[:, :n], 1800, nb)
torch.cuda.synchronize()
ds = E.Dataset.from_device(dev, sp, ex, br, 0, 1800.0, 2000.5)
ch = E.Chains(ds, nch, 1, cfg=E.default_config(0), rep_of_chain=np.arange(nch) % n_rep)
prev = np.zeros((nch, 6)); done = 0
for upto in (2000, 4000, 12000, 22000, 122000):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record(); ch.run_device(upto - done, 0, None); b.record(); torch.cuda.synchronize()
    t = ch.team_stats().astype(float); d = t - prev; prev = t
    print("to %6d: %.2f ms, %.0f ns/it | team iterations: mean %.0f min %.0f max %.0f of %d | hand-overs: total %d, chains %d | commits/it in teams %.4f"
          % (upto, a.elapsed_time(b), a.elapsed_time(b) * 1e6 / (upto - done), d[:, 4].mean(), d[:, 4].min(), d[:, 4].max(), upto - done,
             d[:, 5].sum(), (d[:, 5] > 0).sum(), d[:, 0].sum() / max(d[:, 4].sum(), 1)), flush=True)
    done = upto
