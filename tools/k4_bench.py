"""K4 (direct-over-lineages likelihood, validation path) timing: python tools/k4_bench.py [n_states]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from literate_b200 import engine as E, synth
dev = E.Device(0)
tdev = torch.device("cuda:0")
n, nb = 1_000_000, 200
ts, te = synth.syn_int_device(n, 1, tdev)
ts, te = ts[0, :n].contiguous(), te[0, :n].contiguous()
for ns in [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "8,256,4096").split(",")]:
    lam = torch.rand((ns, nb), dtype=torch.float64, device=tdev) * 0.3 + 0.05
    mu = torch.rand((ns, nb), dtype=torch.float64, device=tdev) * 0.3 + 0.05
    for _ in range(3):
        dev.loglik_direct_device(ts, te, 1800, nb, lam, mu)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        out = dev.loglik_direct_device(ts, te, 1800, nb, lam, mu)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    groups = (ns + 7) // 8
    print("states %d: %.3f ms -> %.3g likelihood evaluations/s over %d lineages; lineage re-reads %.1f GB/s; %.3g lineage-state pairs/s"
          % (ns, ms, ns / (ms * 1e-3), n, 16.0 * n * groups / (ms * 1e-3) / 1e9, ns * n / (ms * 1e-3)))
