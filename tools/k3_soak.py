"""Soak of the speculative team build (loop_variant 4 and the automatic policy) against the compact build: random populations, widths,
leads, launch lengths and sampling periods on three tables; every launch must give the compact build's records, states and counters bit for
bit.  python tools/k3_soak.py [seconds] [seed]   (development aid; run it under `timeout`)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from literate_b200 import engine as E, synth
from oracle import literate_oracle as O

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 300.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
dev = E.Device(0)
tables = {}
ts, te = synth.syn_int(1_000_000); tables["syn-int 1M x 200"] = (ts, te, 0.5)
ts, te = synth.syn_int(3_000, replicate=5); tables["syn-int 3k"] = (ts, te, 0.5)
ts, te = synth.syn_real(200_000, replicate=2); tables["syn-real 200k"] = (ts, te, 0.0)
data = {}
for name, (ts, te, jit) in tables.items():
    st = dev.bin_stats(ts, te, death_jitter=jit, only_dead=True)
    for model in (0, 1, 2, 3):
        data[name + ", model_BDI %d" % model] = (E.Dataset(dev, st, model, float(ts.min()), float(te.max())), model)
t0 = time.time(); n_launch = 0; n_iter_total = 0
while time.time() - t0 < budget:
    name = list(data)[int(rng.integers(len(data)))]
    ds, model = data[name]
    nch = int(rng.choice([1, 2, 37, 148, 149, 256, 296, 297, 592]))
    seed = int(rng.integers(1, 1 << 30))
    W = int(rng.choice([0, 4, 8, 16])); lead = int(rng.integers(1, 8))
    variant = int(rng.choice([0, 4]))
    os.environ.pop("LR_TEAM_W", None); os.environ.pop("LR_TEAM_LEAD", None); os.environ.pop("LR_TEAM_NOBAIL", None)
    if W: os.environ["LR_TEAM_W"] = str(W)
    os.environ["LR_TEAM_LEAD"] = str(lead)
    if rng.uniform() < 0.3: os.environ["LR_TEAM_NOBAIL"] = "1"
    ref = E.Chains(ds, nch, seed=seed, cfg=E.default_config(model, loop_variant=2))
    tst = E.Chains(ds, nch, seed=seed, cfg=E.default_config(model, loop_variant=variant))
    for part in range(int(rng.integers(1, 4))):
        n_it = int(rng.choice([1, 7, 100, 2047, 2048, 2049, 5000, 20000, 60000]))
        se = int(rng.choice([1, 10, 100, 1000])) if n_it <= 5000 else int(rng.choice([100, 1000, 7777]))
        a = ref.run(n_it, se); b = tst.run(n_it, se)
        assert np.array_equal(a, b), ("records differ", name, nch, seed, W, lead, variant, n_it, se, part)
        assert np.array_equal(ref.counters(), tst.counters()), ("counters differ", name, nch, seed, W, lead, variant, n_it, se, part)
        n_launch += 1; n_iter_total += n_it * nch
    ref.close(); tst.close()
    if n_launch % 20 < 3:
        print("%6.0f s  %4d launches  %.3g chain iterations   last: %s, %d chains, W=%d lead=%d variant=%d" % (time.time() - t0, n_launch, n_iter_total, name, nch, W, lead, variant), flush=True)
print("soak ok: %d launches, %.3g chain iterations, all bit-identical to the compact build" % (n_launch, n_iter_total))
