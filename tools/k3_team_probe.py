"""Speculative team build of K3 (loop_variant 4) against the other builds on the bench statistics.
python tools/k3_team_probe.py [chains] [iters]   -- sweeps LR_TEAM_W x LR_TEAM_LEAD, checks the records equal the compact build's."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from literate_b200 import engine as E, synth

nch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
combos = sys.argv[3] if len(sys.argv) > 3 else "8:1,8:2,8:3,8:5,4:2,4:4,16:1,16:2"
dev = E.Device(0)
tdev = torch.device("cuda:0")
n, nb, n_rep = 1_000_000, 200, 4
ts, te = synth.syn_int_device(n, n_rep, tdev)
sp, ex, br = dev.bin_stats_device(ts[:, :n], te[:, :n], 1800, nb)
torch.cuda.synchronize()
ds = E.Dataset.from_device(dev, sp, ex, br, 0, 1800.0, 2000.5)
res = {}

def run(variant, label):
    ch = E.Chains(ds, nch, 1, cfg=E.default_config(0, loop_variant=variant), rep_of_chain=np.arange(nch) % n_rep)
    ch.run(2000)
    nrec = ch.records_per_run(iters, 1000)
    rec = torch.empty((nrec, nch, E.LR_REC_DOUBLES), dtype=torch.float64, device=tdev)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record(); ch.run_device(iters, 1000, rec); b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    res[label] = {"it_per_s": nch * iters / (ms * 1e-3), "ms": ms, "ns_per_it_per_chain": ms * 1e6 / iters}
    t = ch.team_stats().sum(0).astype(float) / (nch * (iters + 2000))
    res[label]["per_iteration"] = {"commits": t[0], "dropped": t[1], "frontier_polls": t[2], "lead_polls": t[3]}
    print(label, res[label], flush=True)
    return rec.cpu().numpy(), ch.counters()

ref, cref = run(2, "compact")
run(0, "auto")
for cmb in combos.split(","):
    W, lead = cmb.split(":")
    os.environ["LR_TEAM_W"] = W; os.environ["LR_TEAM_LEAD"] = lead
    # the library reads the two variables once: reload through a fresh process would be needed -> they are read per call in debug builds
    r, c = run(4, "team_W%s_lead%s" % (W, lead))
    res["team_W%s_lead%s" % (W, lead)]["identical"] = bool(np.array_equal(r, ref) and np.array_equal(c, cref))
    print("   identical to compact:", res["team_W%s_lead%s" % (W, lead)]["identical"], flush=True)
json.dump(res, open("gpurun_out/k3_team_probe.json", "w"), indent=1)
