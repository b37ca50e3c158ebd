"""Random tables through the two builds of K1's real-valued path (k1_bin_kernel / k1_bin_lanes_kernel) and, for small ones, the oracle:
finalized statistics must be identical bit for bit.  python tools/k1_fuzz.py [seconds] [seed]   (development aid)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from literate_b200 import engine as E
from oracle import literate_oracle as O

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
dev = E.Device(0); tdev = torch.device("cuda:0")
t0 = time.time(); cases = 0; lineages = 0
while time.time() - t0 < budget:
    n = int(rng.choice([1, 31, 33, 500, 4097, 70_001, 300_000, 2_000_000]))
    n_rep = int(rng.choice([1, 1, 2, 5])) if n <= 300_000 else 1
    nb = int(rng.choice([1, 2, 24, 37, 200, 311, 439, 440]))
    first = int(rng.choice([-50, 0, 3, 1800, 100_000]))
    pad = int(rng.choice([0, 2, 6]))
    ld = n + pad + ((n + pad) & 1)
    kind = rng.choice(["uniform", "clustered", "sorted", "quarter years", "mostly extant", "outside"])
    span = nb + (6 if kind == "outside" else 0)
    ts = first - (3 if kind == "outside" else 0) + rng.uniform(0, span, (n_rep, n))
    if kind == "clustered": ts = first + np.minimum(rng.exponential(1.5, (n_rep, n)), nb - 1e-9)
    if kind == "quarter years": ts = first + rng.integers(0, 4 * nb, (n_rep, n)) / 4.0
    te = ts + rng.exponential(max(1.0, nb / 5), (n_rep, n))
    if kind == "quarter years": te = ts + rng.integers(0, 4 * nb, (n_rep, n)) / 4.0
    if kind == "mostly extant": te = np.where(rng.uniform(size=(n_rep, n)) < 0.9, first + nb + 0.5, te)
    if kind == "sorted":
        o = np.argsort(ts, axis=1); ts = np.take_along_axis(ts, o, 1); te = np.take_along_axis(te, o, 1)
    if rng.uniform() < 0.2: te[:, : max(1, n // 50)] = ts[:, : max(1, n // 50)]          # zero time at risk
    if rng.uniform() < 0.2: ts[:, -1] = np.nan
    T = torch.full((n_rep, ld), float("nan"), dtype=torch.float64, device=tdev); Tn = torch.full_like(T, float("nan"))
    T[:, :n] = torch.from_numpy(ts).to(tdev); Tn[:, :n] = torch.from_numpy(te).to(tdev)
    fe_ref = float(rng.choice([0.5, 1.0, 0.25]))
    dead = bool(rng.uniform() < 0.3); end_time = float(first + nb * rng.uniform(0.5, 1.1))
    out = {}
    for build in ("0", "1"):
        os.environ["LR_K1_LANES"] = build
        out[build] = [x.cpu().numpy() for x in dev.bin_stats_device(T[:, :n], Tn[:, :n], first, nb, fe_ref=fe_ref, dead_only=dead, end_time=end_time)]
    for a, b, what in zip(out["0"], out["1"], ("sp", "ex", "br")):
        assert np.array_equal(a, b, equal_nan=True), (what, n, n_rep, nb, first, kind, fe_ref, dead)
    if n <= 4097 and nb <= 37 and not dead:
        with np.errstate(invalid="ignore"):
            for r in range(n_rep):
                for j in (0, nb // 2, nb - 1):
                    a, b, c = O.events_in_bin(ts[r], te[r], first + j, first + j + 1)
                    assert (a, b) == (out["1"][0][r, j], out["1"][1][r, j]) and abs(c - out["1"][2][r, j]) <= 1e-12 * max(1.0, abs(c)), (n, nb, first, kind, j)
    cases += 1; lineages += n * n_rep
os.environ.pop("LR_K1_LANES", None)
print("k1 fuzz ok: %d tables, %.3g lineages, the two builds identical bit for bit" % (cases, lineages))
