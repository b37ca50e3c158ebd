"""engine.Pipeline (host tables in, records out, K1 of batch k+1 beside the chains of batch k) against the plain sequence on a separate
handle: random batch sizes, fp64 and int32 tables, several staging batches per table.  python tools/pipe_soak.py [seconds] [seed]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from literate_b200 import engine as E, synth

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 90.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
pipe = E.Pipeline(0); ref_dev = E.Device(0)
t0 = time.time(); pushed = []; checked = 0

def expect(job):
    ts, te, kw = job
    st = ref_dev.bin_stats(ts, te, first_bin=kw["first_bin"], n_bins=kw["n_bins"], death_jitter=kw["death_jitter"], end_time=kw["end_time"])
    ds = E.Dataset(ref_dev, st, 0, kw["start_time"], kw["end_time"])
    ch = E.Chains(ds, kw["n_chains"], kw["seed"], E.default_config(0), rep_of_chain=kw["rep_of_chain"])
    rec = ch.run(kw["n_iter"], kw["sample_every"])
    ch.close(); ds.close()
    return rec, st

def check(done, job):
    global checked
    rec, st = done
    erec, est = expect(job)
    assert np.array_equal(rec, erec), "records differ"
    assert (st.sp == est.sp).all() and (st.ex == est.ex).all() and (st.br == est.br).all(), "statistics differ"
    checked += 1

while time.time() - t0 < budget:
    n = int(rng.choice([5_000, 200_000, 1_000_000])); n_rep = int(rng.choice([1, 3, 8]))
    ts = np.empty((n_rep, n)); te = np.empty((n_rep, n))
    for r in range(n_rep):
        ts[r], te[r] = synth.syn_int(n, replicate=int(rng.integers(1 << 20)))
    kw = dict(n_chains=n_rep * int(rng.choice([1, 2])), n_iter=int(rng.choice([300, 3000, 20000])), sample_every=int(rng.choice([100, 1000])),
              seed=int(rng.integers(1, 1 << 30)), first_bin=1800, n_bins=200, start_time=1800.0, end_time=2000.5, death_jitter=0.5)
    kw["rep_of_chain"] = (np.arange(kw["n_chains"]) % n_rep).astype(np.int32)
    if rng.uniform() < 0.4:
        ts, te = ts.astype(np.int32), (te - 0.5).astype(np.int32)
    job = (ts, te, kw)
    done = pipe.push(ts, te, kw["n_chains"], kw["n_iter"], kw["sample_every"], seed=kw["seed"], first_bin=1800, n_bins=200, death_jitter=0.5,
                     start_time=1800.0, end_time=2000.5, rep_of_chain=kw["rep_of_chain"])
    if done is not None:
        check(done, pushed[-1])
    pushed.append(job)
    pushed = pushed[-1:]
check(pipe.flush(), pushed[-1])
print("pipeline soak ok: %d batches, records and statistics equal the plain sequence's" % checked)
