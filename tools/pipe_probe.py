"""Phase timings of engine.Pipeline (development aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from literate_b200 import engine as E, synth
tdev = torch.device("cuda:0")
n, n_rep, chains, iters = 1_000_000, 256, 256, int(os.environ.get('PROBE_ITERS', 100_000))
ts, te = synth.syn_int_device(n, n_rep, tdev)
hts = torch.empty((n_rep, n), dtype=torch.float64, pin_memory=True); hts.copy_(ts[:, :n])
hte = torch.empty((n_rep, n), dtype=torch.float64, pin_memory=True); hte.copy_(te[:, :n])
hrec = torch.empty((100, chains, 144), dtype=torch.float64, pin_memory=True)
nts, nte = hts.numpy(), hte.numpy()
rep = np.arange(chains, dtype=np.int32)
pipe = E.Pipeline(0)
T = time.perf_counter
orig_bin = pipe.dev_bin.bin_stats
def timed_bin(*a, **k):
    t = T(); r = orig_bin(*a, **k); print("   bin %.1f ms" % ((T() - t) * 1e3)); return r
pipe.dev_bin.bin_stats = timed_bin
orig_collect = pipe._collect
def timed_collect(out):
    t = T(); r = orig_collect(out); print("   collect %.1f ms" % ((T() - t) * 1e3)); return r
pipe._collect = timed_collect
for k in range(6):
    t0 = T()
    pipe.push(nts, nte, chains, iters, 1000, seed=k, first_bin=1800, n_bins=200, start_time=1800.0, end_time=2000.5, rep_of_chain=rep, out=hrec)
    print("push %d: %.1f ms" % (k, (T() - t0) * 1e3), flush=True)
t0 = T(); pipe.flush(out=hrec); print("flush %.1f ms" % ((T() - t0) * 1e3))
t0 = T(); torch.cuda.synchronize(); print("device sync after flush %.1f ms" % ((T() - t0) * 1e3))
for k in range(0):
    pipe.push(nts, nte, chains, iters, 1000, seed=k, first_bin=1800, n_bins=200, start_time=1800.0, end_time=2000.5, rep_of_chain=rep, out=hrec)
t0 = T(); pipe.dev_run.sync(); print("dev_run sync %.1f ms" % ((T() - t0) * 1e3))
t0 = T(); pipe.dev_bin.sync(); print("dev_bin sync %.1f ms" % ((T() - t0) * 1e3))
t0 = T(); torch.cuda.current_stream().synchronize(); print("torch stream sync %.1f ms" % ((T() - t0) * 1e3))
t0 = T(); torch.cuda.synchronize(); print("device sync %.1f ms" % ((T() - t0) * 1e3))
