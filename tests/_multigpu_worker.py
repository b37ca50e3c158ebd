"""Multi-GPU parity worker: run as `python -m torch.distributed.run --nproc-per-node N tests/_multigpu_worker.py`
(tests/test_gpu_multi.py does, when >= 2 GPUs are visible).  Every check compares the sharded path with the
single-device path computed on the same rank."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from literate_b200 import engine as E, parallel as P, synth


def main():
    rank, local, world = P.init()
    assert world >= 2 and dist.get_backend() == "nccl"
    tdev = torch.device("cuda", local)
    dev = E.Device(local)

    # ---- 1. lineage-sharded binning + int64 all-reduce == whole-table binning, bit for bit
    for real in (0, 1):
        n = 400_003
        ts, te = synth.syn_real(n, replicate=5) if real else synth.syn_int(n, replicate=5)
        fe_ref = 1.0 if real else 0.5
        s0, cnt = P.shard_range(n, world, rank)
        lts = torch.from_numpy(ts[s0:s0 + cnt]).to(tdev); lte = torch.from_numpy(te[s0:s0 + cnt]).to(tdev)
        first, nb, lo, hi = P.global_window(lts, lte)
        assert (first, nb, lo, hi) == (int(ts.min()), int(te.max()) - int(ts.min()), ts.min(), te.max())
        sp, ex, br = P.bin_stats_lineage_sharded(dev, lts, lte, first, nb, fe_ref=fe_ref)
        fts, fte = torch.from_numpy(ts).to(tdev), torch.from_numpy(te).to(tdev)
        sp1, ex1, br1 = dev.bin_stats_device(fts, fte, first, nb, fe_ref=fe_ref)
        torch.cuda.synchronize()
        assert torch.equal(sp, sp1) and torch.equal(ex, ex1) and torch.equal(br, br1), ("lineage shards differ", real)
        assert int(sp.sum()) == n

    # ---- 2. chain sharding: a shard reproduces its slice of the whole population
    sp, ex, br = sp1, ex1, br1
    ds = E.Dataset.from_device(dev, sp, ex, br, 0, float(ts.min()), float(te.max()))
    n_chains = 8 * world
    c0, nl = P.shard_range(n_chains, world, rank)
    mine = E.Chains(ds, nl, seed=9, chain_id0=c0)
    whole = E.Chains(ds, n_chains, seed=9)
    rm, rw = mine.run(3000, 100), whole.run(3000, 100)
    assert np.array_equal(rm, rw[:, c0:c0 + nl]), "chain shard differs from the population"

    # ---- 3. a tempered ladder that spans all ranks: gathered (lik, beta) table, same decisions everywhere
    T = n_chains
    beta = P.temperature_ladder(T, 0.2)
    mine.set_beta(beta[c0:c0 + nl]); whole.set_beta(beta)
    for rnd in range(6):
        table = P.tempered_swap(mine, c0, T, rnd)
        whole.swap_step(T, rnd)
        torch.cuda.synchronize()
        assert table.shape == (n_chains, 2)
        got, want = mine.state()[:, E.REC_BETA], whole.state()[:, E.REC_BETA]
        assert np.array_equal(got, want[c0:c0 + nl]), ("swap round differs", rnd)
        mine.run(400); whole.run(400)
    assert np.array_equal(mine.state(), whole.state()[c0:c0 + nl])
    acc = whole.counters()[:, 9].sum()

    # ---- 4. the sibling samplers: a chain shard reproduces its slice of the population
    from literate_b200 import trend as TR, ddrate as DD
    hsp, hex_, hbr = sp.cpu().numpy()[0], ex.cpu().numpy()[0], br.cpu().numpy()[0]
    nbin = hsp.shape[0]
    trend = np.clip(np.linspace(0.0, 1.0, nbin), 1e-15, 1.0)
    t_mine = TR.TrendChains(dev, hsp, hex_, hbr, trend, nl, 9, chain_id0=c0)
    t_all = TR.TrendChains(dev, hsp, hex_, hbr, trend, n_chains, 9)
    assert np.array_equal(t_mine.run(2001, 100), t_all.run(2001, 100)[:, c0:c0 + nl]), "TrendRate shard differs"
    d_mine = DD.DDChains(dev, hsp, hex_, hbr, float(ts.min()), float(te.max()), 2, 2, None, None, nl, 9, chain_id0=c0)
    d_all = DD.DDChains(dev, hsp, hex_, hbr, float(ts.min()), float(te.max()), 2, 2, None, None, n_chains, 9)
    assert np.array_equal(d_mine.run(2001, 100), d_all.run(2001, 100)[:, c0:c0 + nl]), "DDRate shard differs"
    dist.barrier()
    if rank == 0:
        print("MULTIGPU OK world=%d swaps_accepted=%d" % (world, acc), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
