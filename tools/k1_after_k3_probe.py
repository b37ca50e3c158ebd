"""K1 timed alone, after an idle gap, and after a K3 launch (why bench.py's K1 is 5 % slower than tools/k1_bench.py's): development aid."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from literate_b200 import engine as E, synth

dev = E.Device(0); tdev = torch.device("cuda:0")
n, nb, n_rep, chains = 1_000_000, 200, 256, 256
ts, te = synth.syn_int_device(n, n_rep, tdev)
ts, te = ts[:, :n], te[:, :n]
acc = dev.new_accumulators(n_rep, nb, tdev)
sp, ex, br = dev.bin_stats_device(ts, te, 1800, nb)
torch.cuda.synchronize()
ds = E.Dataset.from_device(dev, sp, ex, br, 0, 1800.0, 2000.5)
rec = torch.empty((100, chains, E.LR_REC_DOUBLES), dtype=torch.float64, device=tdev)

def k1_timed():
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    acc.zero_()
    a.record(); dev.bin_accumulate_device(ts, te, 1800, nb, acc, fe_ref=0.5); b.record()
    return a, b

def run(label, between):
    evs = []
    for k in range(8):
        between(k)
        evs.append(k1_timed())
    torch.cuda.synchronize()
    ms = [a.elapsed_time(b) for a, b in evs][2:]
    print("%-34s K1 %.4f ms (min %.4f max %.4f) = %.0f GB/s" % (label, statistics.mean(ms), min(ms), max(ms), 16.0 * n * n_rep / statistics.mean(ms) / 1e6), flush=True)

run("back to back", lambda k: None)
run("after a host sync", lambda k: torch.cuda.synchronize())
run("after 25 ms of torch._sleep", lambda k: torch.cuda._sleep(int(25e-3 * 1.9e9)))
def k3(k, variant=0):
    ch = E.Chains(ds, chains, seed=k + 1, cfg=E.default_config(0, loop_variant=variant), rep_of_chain=np.arange(chains))
    ch.run_device(100_000, 1000, rec)
    k3.keep.append(ch)
k3.keep = []
run("after K3 (teams, 25 ms)", lambda k: k3(k))
run("after K3 (compact build, 70 ms)", lambda k: k3(k, 2))
big = torch.empty(1 << 30, dtype=torch.uint8, device=tdev)
run("after a 1 GiB memset", lambda k: big.zero_())

def k3_then_sleep(k):
    k3(k); torch.cuda._sleep(int(25e-3 * 1.9e9))
run("after K3 (teams) + 25 ms idle", k3_then_sleep)
def k3_then_k1(k):
    k3(k); acc.zero_(); dev.bin_accumulate_device(ts, te, 1800, nb, acc, fe_ref=0.5)
run("second K1 after K3 (teams)", k3_then_k1)
def k3_short(k):
    ch = E.Chains(ds, chains, seed=k + 1, rep_of_chain=np.arange(chains)); ch.run_device(10_000, 1000, rec[:10]); k3.keep.append(ch)
run("after K3 (teams, 10 000 iterations)", k3_short)
def k3_w4(k):
    os.environ["LR_TEAM_W"] = "4"; k3(k); os.environ.pop("LR_TEAM_W")
run("after K3 (teams of 4)", k3_w4)
