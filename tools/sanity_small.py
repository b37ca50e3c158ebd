"""Smallest end-to-end exercise of every kernel (for compute-sanitizer runs; development aid)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from literate_b200 import engine as E, synth, parallel as P

dev = E.Device(0)
for real in (0, 1):
    ts, te = synth.syn_real(3001) if real else synth.syn_int(3001)
    st = dev.bin_stats(np.stack([ts, ts]), np.stack([te, te]), only_dead=True, death_jitter=0.0 if real else 0.5)
    assert st.sp.sum() == 2 * 3001
ts, te = synth.syn_int(3001)
st = dev.bin_stats(ts, te, only_dead=True)
for model in (0, 3):
    ds = E.Dataset(dev, st, model, float(ts.min()), float(te.max()))
    for variant in (1, 2):
        ch = E.Chains(ds, 12, seed=3, cfg=E.default_config(model, loop_variant=variant))
        rec = ch.run(1501, 250)
        assert np.isfinite(rec[:, :, E.REC_LIK]).all()
        ch.set_beta(np.tile(P.temperature_ladder(4, 0.2), 3))
        for r in range(3):
            ch.swap_step(4, r); ch.run(100)
        out = ds.evaluate([E.record_to_state(x, float(te.max())) for x in rec[-1]])
        s = ch.state(); ch.set_state(s)
        ch.close()
    ds.close()
print("sanity ok, launches", dev.kernel_launches)
