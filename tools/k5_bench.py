"""K5 timing (development aid): 4096 chains x 500 samples of real chain output, 2.36 GB of records."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from literate_b200 import engine as E, synth
dev = E.Device(0); tdev = torch.device("cuda:0")
ts, te = synth.syn_int_device(1_000_000, 1, tdev)
sp, ex, br = dev.bin_stats_device(ts[:, :1_000_000], te[:, :1_000_000], 1800, 200)
torch.cuda.synchronize()
ds = E.Dataset.from_device(dev, sp, ex, br, 0, 1800.0, 2000.5)
ch = E.Chains(ds, 4096, 1)
ch.run(3000)
rec = torch.empty((500, 4096, 144), dtype=torch.float64, device=tdev)
ch.run_device(50000, 100, rec, stream="handle"); dev.sync()
for _ in range(2):
    dev.summarize_records_device(rec, 1800.0, 200)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    out = dev.summarize_records_device(rec, 1800.0, 200)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 5
n = rec.shape[0] * rec.shape[1]
print("K5: %d records (%.2f GB) in %.3f ms = %.3g records/s, %.0f GB/s of records" % (n, n * 1152 / 1e9, ms, n / (ms * 1e-3), n * 1152 / (ms * 1e-3) / 1e9))
# HPD intervals on the device: the per-sample matrix a range of bins at a time (lr_marginal_rates) + torch.sort of its columns
from literate_b200 import summary as S
S.summarize_records_device(dev, rec, 1800.0, 2000.5, burnin=0.0, hpd=True)
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
out = S.summarize_records_device(dev, rec, 1800.0, 2000.5, burnin=0.0, hpd=True)
torch.cuda.synchronize()
print("means + HPD of birth, death, net rate for %d records x 200 bins on the device: %.1f ms" % (n, 1e3 * (time.perf_counter() - t0)))
a.record(); mb, md = dev.marginal_rates_device(rec, 1800.0, 200, 0, 64); b.record(); torch.cuda.synchronize()
print("lr_marginal_rates, 64 bins: %.3f ms (%.0f GB/s written)" % (a.elapsed_time(b), 2 * n * 64 * 8 / (a.elapsed_time(b) * 1e-3) / 1e9))
