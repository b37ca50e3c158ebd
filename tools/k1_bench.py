"""K1 timing probe: every variant on integer / real / sorted tables (development aid; bench.py is the measured entry point).
   python tools/k1_bench.py [n_rep] [variants, e.g. 2,3] [kinds, e.g. int,real,sorted] [reps]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from literate_b200 import engine as E, synth

n_rep = int(sys.argv[1]) if len(sys.argv) > 1 else 64
variants = [0]   # a single kernel is left; the argument is kept for old command lines
kinds = (sys.argv[3] if len(sys.argv) > 3 else "int,real,sorted").split(",")
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 10
dev = E.Device(0)
tdev = torch.device("cuda:0")
n, nb = int(os.environ.get('K1_BENCH_N', '1000000')), int(os.environ.get('K1_BENCH_NB', '200'))
res = {}
for kind in kinds:
    if kind in ("real", "realsorted"):
        ts, te = synth.syn_real_device(n, n_rep, tdev); fe = 1.0
        if kind == "realsorted":
            ts, idx = torch.sort(ts, dim=1); te = torch.gather(te, 1, idx)
    else:
        ts, te = synth.syn_int_device(n, n_rep, tdev); fe = 0.5
        if kind == "sorted":
            ts, idx = torch.sort(ts, dim=1); te = torch.gather(te, 1, idx)
    ts, te = ts[:, :n], te[:, :n]
    if nb != 200:                      # stretch the 200-year window of the synthetic tables over nb bins
        ts = 1800.0 + (ts - 1800.0) * (nb / 200.0); te = 1800.0 + (te - 1800.0) * (nb / 200.0)
    ref = None
    for variant in variants:
        acc = dev.new_accumulators(n_rep, nb, tdev)
        dev.bin_accumulate_device(ts, te, 1800, nb, acc, fe_ref=fe)
        out = [t.clone() for t in dev.bin_finalize_device(acc, nb, fe_ref=fe)]
        torch.cuda.synchronize()
        if ref is None:
            ref = out
        ok = all(bool((a == b).all()) for a, b in zip(ref, out))
        for force in ("0", "1"):      # the two builds of the real-valued path against each other
            os.environ["LR_K1_LANES"] = force
            acc2 = dev.new_accumulators(n_rep, nb, tdev)
            dev.bin_accumulate_device(ts, te, 1800, nb, acc2, fe_ref=fe)
            out2 = dev.bin_finalize_device(acc2, nb, fe_ref=fe)
            torch.cuda.synchronize()
            ok = ok and all(bool((a == b).all()) for a, b in zip(ref, out2))
        os.environ.pop("LR_K1_LANES")
        if len(sys.argv) > 5: os.environ["LR_K1_LANES"] = sys.argv[5]
        for _ in range(2):
            dev.bin_accumulate_device(ts, te, 1800, nb, acc, fe_ref=fe)
        torch.cuda.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        for a, b in evs:
            a.record(); dev.bin_accumulate_device(ts, te, 1800, nb, acc, fe_ref=fe); b.record()
        torch.cuda.synchronize()
        ms = sorted(a.elapsed_time(b) for a, b in evs)
        gbs = 16.0 * n * n_rep / (ms[len(ms) // 2] * 1e-3) / 1e9
        res[f"k1_{kind}_v{variant}"] = {"ms_med": ms[len(ms) // 2], "ms_min": ms[0], "GBps": gbs, "same_as_first_variant": ok}
        print(kind, variant, res[f"k1_{kind}_v{variant}"], flush=True)
    del ts, te
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/k1_bench.json", "w"), indent=1)
