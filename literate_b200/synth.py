"""Deterministic synthetic lineage tables.

syn_int : integer years, ts = 1800 + floor(200 U), te = min(ts + floor(Exp(8)), 2000); with the default
          -death_jitter .5 every value is a half-integer (the shape of all data the reference ships).
syn_real: real-valued times, ts = 1800 + 200 U, te = min(ts + Exp(8), 2000), for -death_jitter 0.
Lineage 0 is forced to (1800, 2000) so the window is exactly 1800..2000 (200 unit bins).
"""
from __future__ import annotations

import numpy as np

T0, SPAN, MEAN_LIFE = 1800, 200, 8.0
BASE_SEED = 20260101


def syn_int(n, replicate=0, jitter=0.5):
    g = np.random.Generator(np.random.Philox(BASE_SEED + replicate))
    ts = T0 + np.floor(SPAN * g.random(n))
    te = np.minimum(ts + np.floor(g.exponential(MEAN_LIFE, n)), T0 + SPAN)
    ts[0], te[0] = T0, T0 + SPAN
    return ts, te + jitter


def syn_real(n, replicate=0):
    g = np.random.Generator(np.random.Philox(BASE_SEED + 7919 + replicate))
    ts = T0 + SPAN * g.random(n)
    te = np.minimum(ts + g.exponential(MEAN_LIFE, n), T0 + SPAN)
    ts[0], te[0] = T0, T0 + SPAN
    return ts, te


def syn_int_device(n, n_rep, device, jitter=0.5, seed=BASE_SEED):
    """Same distribution generated on the GPU with torch (bench inputs at sizes the host RNG is too slow for).
    Returns float64 CUDA tensors [n_rep, ld] with ld = n rounded up to even; use [:, :n]."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    ld = (n + 1) & ~1
    ts = torch.empty((n_rep, ld), dtype=torch.float64, device=device)
    te = torch.empty((n_rep, ld), dtype=torch.float64, device=device)
    for r in range(n_rep):
        u = torch.rand(ld, generator=g, device=device, dtype=torch.float64)
        s = T0 + torch.floor(SPAN * u)
        life = torch.floor(-MEAN_LIFE * torch.log1p(-torch.rand(ld, generator=g, device=device, dtype=torch.float64)))
        e = torch.minimum(s + life, torch.tensor(float(T0 + SPAN), device=device, dtype=torch.float64)) + jitter
        s[0], e[0] = T0, T0 + SPAN + jitter
        ts[r], te[r] = s, e
    return ts, te


def syn_real_device(n, n_rep, device, seed=BASE_SEED + 7919):
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    ld = (n + 1) & ~1
    ts = torch.empty((n_rep, ld), dtype=torch.float64, device=device)
    te = torch.empty((n_rep, ld), dtype=torch.float64, device=device)
    for r in range(n_rep):
        s = T0 + SPAN * torch.rand(ld, generator=g, device=device, dtype=torch.float64)
        life = -MEAN_LIFE * torch.log1p(-torch.rand(ld, generator=g, device=device, dtype=torch.float64))
        e = torch.minimum(s + life, torch.tensor(float(T0 + SPAN), device=device, dtype=torch.float64))
        s[0], e[0] = T0, T0 + SPAN
        ts[r], te[r] = s, e
    return ts, te
