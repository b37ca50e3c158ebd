"""Quick K1/K3 timing probe (development aid; bench.py is the measured entry point)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from literate_b200 import engine as E, synth

dev = E.Device(0)
tdev = torch.device("cuda:0")
n, n_rep, nb = 1_000_000, int(os.environ.get("NREP", 64)), 200
res = {}
for kind in ("int", "real", "sorted"):
    if kind == "int":
        ts, te = synth.syn_int_device(n, n_rep, tdev); fe = 0.5
    elif kind == "real":
        ts, te = synth.syn_real_device(n, n_rep, tdev); fe = 1.0
    else:
        ts, te = synth.syn_int_device(n, n_rep, tdev); fe = 0.5
        ts, idx = torch.sort(ts, dim=1); te = torch.gather(te, 1, idx)
    ts, te = ts[:, :n], te[:, :n]
    for variant in (1, 2):
        dev.set_bin_kernel(variant)
        acc = dev.new_accumulators(n_rep, nb, tdev)
        for _ in range(3):
            dev.bin_accumulate_device(ts, te, 1800, nb, acc, fe_ref=fe)
        torch.cuda.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
        for a, b in evs:
            a.record(); dev.bin_accumulate_device(ts, te, 1800, nb, acc, fe_ref=fe); b.record()
        torch.cuda.synchronize()
        ms = sorted(a.elapsed_time(b) for a, b in evs)
        gbs = 16.0 * n * n_rep / (ms[len(ms) // 2] * 1e-3) / 1e9
        res[f"k1_{kind}_v{variant}"] = {"ms_med": ms[len(ms) // 2], "ms_min": ms[0], "GBps": gbs}
        print(kind, variant, res[f"k1_{kind}_v{variant}"], flush=True)
    del ts, te
dev.set_bin_kernel(0)

# K3 probe
ts, te = synth.syn_int_device(n, 4, tdev)
sp, ex, br = dev.bin_stats_device(ts[:, :n], te[:, :n], 1800, nb)
torch.cuda.synchronize()
ds = E.Dataset.from_device(dev, sp, ex, br, 0, 1800.0, 2000.5)
for nch in (256, 4096):
    ch = E.Chains(ds, nch, 1, rep_of_chain=np.arange(nch) % 4)
    ch.run(2000)
    for iters in (20000,):
        t0 = time.time(); ch.run(iters); dt = time.time() - t0
        cnt = ch.counters().sum(0)
        res[f"k3_{nch}"] = {"it_per_s": nch * iters / dt, "s": dt, "acc_rate": float(cnt[1] / cnt[0])}
        print(nch, res[f"k3_{nch}"], flush=True)
    st = ch.state()
    print("K_l mean", st[:, E.REC_KL].mean(), "K_m mean", st[:, E.REC_KM].mean(), "lik", st[:, E.REC_LIK].mean())
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/k1_probe.json", "w"), indent=1)
