"""Does a running K3 launch slow down host->device copies or K1?  (development aid for the end-to-end pipeline)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from literate_b200 import engine as E, synth
tdev = torch.device("cuda:0")
dev = E.Device(0)
n, nb, n_rep = 1_000_000, 200, 64
ts, te = synth.syn_int_device(n, n_rep, tdev)
ts, te = ts[:, :n], te[:, :n]
sp, ex, br = dev.bin_stats_device(ts, te, 1800, nb)
torch.cuda.synchronize()
ds = E.Dataset.from_device(dev, sp, ex, br, 0, 1800.0, 2000.5)
host = torch.empty(1 << 30, dtype=torch.uint8, pin_memory=True); host.fill_(1)
devb = torch.empty(1 << 30, dtype=torch.uint8, device=tdev)
acc = dev.new_accumulators(n_rep, nb, tdev)
side = torch.cuda.Stream()
for variant in (1, 4):
    ch = E.Chains(ds, 256, 1, cfg=E.default_config(0, loop_variant=variant), rep_of_chain=np.arange(256) % n_rep)
    ch.run(6000)
    for busy in (False, True):
        torch.cuda.synchronize()
        if busy:
            ch.run_device(600000, 0, None)          # ~150-250 ms of K3 on torch's current stream
        a, b, c, d = (torch.cuda.Event(enable_timing=True) for _ in range(4))
        with torch.cuda.stream(side):
            a.record()
            for _ in range(4):
                devb.copy_(host, non_blocking=True)
            b.record()
            c.record()
            for _ in range(10):
                dev.bin_accumulate_device(ts, te, 1800, nb, acc, fe_ref=0.5, stream=side.cuda_stream)
            d.record()
        torch.cuda.synchronize()
        print("K3 variant %d %s: H2D %.1f GB/s, K1 (64 replicates) %.3f ms per launch = %.0f GB/s"
              % (variant, "RUNNING" if busy else "idle", 4 * (1 << 30) / (a.elapsed_time(b) * 1e-3) / 1e9, c.elapsed_time(d) / 10,
                 16.0 * n * n_rep / (c.elapsed_time(d) / 10 * 1e-3) / 1e9), flush=True)
    ch.close()
