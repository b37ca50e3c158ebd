"""How often does an iteration CHANGE the chain's state?  (the quantity that bounds speculative evaluation in K3)
python tools/accept_probe.py [chains] -- cfg3 statistics, counters after 2k / 10k / 100k iterations from the initial state."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from literate_b200 import engine as E, synth

nch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = E.Device(0)
tdev = torch.device("cuda:0")
n, nb, n_rep = 1_000_000, 200, 4
ts, te = synth.syn_int_device(n, n_rep, tdev)
sp, ex, br = dev.bin_stats_device(ts[:, :n], te[:, :n], 1800, nb)
torch.cuda.synchronize()
ds = E.Dataset.from_device(dev, sp, ex, br, 0, 1800.0, 2000.5)
ch = E.Chains(ds, nch, 1, cfg=E.default_config(0), rep_of_chain=np.arange(nch) % n_rep)
out = {}
prev = np.zeros(10)
done = 0
for upto in (500, 2000, 10000, 30000, 100000, 200000):
    ch.run(upto - done); done = upto
    cnt = ch.counters().sum(0).astype(float)
    d = cnt - prev; prev = cnt
    st = ch.state()
    out[str(upto)] = {"iters": d[0], "accepted": d[1] / d[0], "noop_moves": d[4] / d[0], "state_changes": (d[1] - d[4]) / d[0],
                      "rate_props": d[3] / d[0], "rj": d[5] / d[0], "gibbs": d[6] / d[0],
                      "K_l": float(st[:, E.REC_KL].mean()), "K_m": float(st[:, E.REC_KM].mean())}
    print(upto, out[str(upto)], flush=True)
json.dump(out, open("gpurun_out/accept_probe.json", "w"), indent=1)
