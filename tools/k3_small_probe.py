"""K3 builds on the SMALL tables (high acceptance: many state changes per iteration), where speculation pays least.
python tools/k3_small_probe.py [chains] [iters]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from literate_b200 import engine as E
from oracle import literate_oracle as O      # development tool: parsing only

nch = int(sys.argv[1]) if len(sys.argv) > 1 else 64
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 50000
combos = sys.argv[3] if len(sys.argv) > 3 else "8:1,8:2,8:4,16:1,16:2,16:4,4:2"
dev = E.Device(0)
res = {}
for name in ("example_dataTAD.txt", "metal_bands_1.tsv"):
    src = os.path.join("tests", "golden", "inputs", name)
    if not os.path.exists(src):
        import gzip, shutil
        with gzip.open(src + ".gz", "rb") as fi, open("/tmp/" + name, "wb") as fo:
            shutil.copyfileobj(fi, fo)
        src = "/tmp/" + name
    lin = O.read_lineages(src)
    st = dev.bin_stats(lin.ts, lin.te)
    ds = E.Dataset(dev, st, 0, lin.start_time, lin.end_time)
    def run(variant, label):
        ch = E.Chains(ds, nch, 1, cfg=E.default_config(0, loop_variant=variant))
        ch.run(5000)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record(); ch.run_device(iters, 0, None); b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        cnt = ch.counters().sum(0).astype(float)
        t = ch.team_stats().sum(0).astype(float) / cnt[0]
        res[name + " " + label] = {"it_per_s": nch * iters / (ms * 1e-3), "ns_per_it_per_chain": ms * 1e6 / iters,
                                   "accepted": cnt[1] / cnt[0], "noop_moves": cnt[4] / cnt[0], "commits": t[0], "dropped": t[1], "lead_polls": t[3]}
        print(name, label, res[name + " " + label], flush=True)
    run(1, "spec1"); run(2, "compact")
    for cmb in combos.split(","):
        W, lead = cmb.split(":")
        os.environ["LR_TEAM_W"] = W; os.environ["LR_TEAM_LEAD"] = lead
        run(4, "team_W%s_lead%s" % (W, lead))
json.dump(res, open("gpurun_out/k3_small_probe.json", "w"), indent=1)
