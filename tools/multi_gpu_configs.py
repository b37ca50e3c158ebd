"""BASELINE configs[3] and configs[4] on N GPUs (run under torchrun, one rank per GPU; development/measurement aid):

  cfg4  synthetic 1M lineages, 4096 chains (replicates + tempered swaps) sharded over the ranks:
        512 ladders of 8 temperatures, ladders never span ranks -> no collective in the loop; plus the cross-rank variant
        (one swap all-gather per round) for comparison.
  cfg5  synthetic 100M lineages with the lineage axis sharded over the ranks, per-bin sufficient statistics combined by
        one int64 all-reduce over NCCL.

Prints one JSON object on rank 0 (timings are CUDA events, max over ranks).
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from literate_b200 import engine as E, parallel as P, synth


def tmax(x, tdev):
    t = torch.tensor([x], dtype=torch.float64, device=tdev)
    if dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def main():
    rank, local, world = P.init()
    torch.cuda.set_device(local)
    tdev = torch.device("cuda", local)
    dev = E.Device(local)
    out = {"world": world}
    nb, first = 200, 1800

    # ------------------------------------------------------------------ cfg5: lineage-sharded binning
    n_total = int(os.environ.get("LR_CFG5_LINEAGES", 100_000_000))
    s0, cnt = P.shard_range(n_total, world, rank)
    ts, te = synth.syn_int_device(cnt, 1, tdev, seed=synth.BASE_SEED + 17 * rank)
    ts, te = ts[:, :cnt], te[:, :cnt]
    if rank != 0:                       # lineage 0 of every shard was forced to span the window; harmless, but keep totals simple
        pass
    for _ in range(3):
        sp, ex, br = P.bin_stats_lineage_sharded(dev, ts, te, first, nb, fe_ref=0.5)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    evs = []
    for _ in range(10):
        a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        acc = dev.new_accumulators(1, nb, tdev)
        a.record()
        dev.bin_accumulate_device(ts, te, first, nb, acc, fe_ref=0.5)
        b.record()
        P.allreduce_accumulators(acc)
        sp, ex, br = dev.bin_finalize_device(acc, nb, fe_ref=0.5)
        c.record()
        evs.append((a, b, c))
    torch.cuda.synchronize()
    k1 = sorted(a.elapsed_time(b) for a, b, c in evs)[5]
    tot = sorted(a.elapsed_time(c) for a, b, c in evs)[5]
    k1, tot = tmax(k1, tdev), tmax(tot, tdev)
    assert int(sp.sum()) == n_total, (int(sp.sum()), n_total)                 # every lineage is born once inside the window
    want_br = (torch.minimum(te, torch.tensor(2000.0, device=tdev, dtype=torch.float64)) - ts).sum()
    tb = want_br.reshape(1).clone()
    if world > 1:
        dist.all_reduce(tb)
    assert float(br.sum()) == float(tb[0]), (float(br.sum()), float(tb[0]))   # total time at risk conserved (half-integers: exact)
    out["cfg5"] = {"lineages": n_total, "per_gpu": cnt, "k1_ms": k1, "k1_plus_allreduce_finalize_ms": tot,
                   "aggregate_GBps": 16.0 * n_total / (k1 * 1e-3) / 1e9, "aggregate_GBps_incl_allreduce": 16.0 * n_total / (tot * 1e-3) / 1e9,
                   "allreduce_bytes": int(acc.numel() * 8)}
    del ts, te
    if rank == 0:
        print(json.dumps({"cfg5": out["cfg5"]}), file=sys.stderr, flush=True)

    # ------------------------------------------------------------------ cfg4: 4096 chains, tempered, sharded
    n_chains, T, iters, swap_every = int(os.environ.get("LR_CFG4_CHAINS", 4096)), 8, 100_000, 1000
    ts, te = synth.syn_int_device(1_000_000, 1, tdev)
    sp, ex, br = dev.bin_stats_device(ts[:, :1_000_000], te[:, :1_000_000], first, nb)
    torch.cuda.synchronize()
    ds = E.Dataset.from_device(dev, sp, ex, br, 0, 1800.0, 2000.5)
    for mode in ("ladders_local", "ladders_span_ranks"):
        if mode == "ladders_local":
            ladder, total = T, n_chains
        else:
            # ladders of 32 consecutive chain ids with 16.5 ladders per rank: every rank boundary cuts a ladder in two,
            # so each swap round needs the all-gather of the (lik, beta) table
            ladder = 32
            total = (n_chains // world // ladder * ladder + ladder // 2) * world if world > 1 else n_chains
        c0, nl = P.shard_range(total, world, rank)
        ch = E.Chains(ds, nl, seed=77, chain_id0=c0)
        beta = P.temperature_ladder(ladder, 0.1 * T / ladder)[np.arange(c0, c0 + nl) % ladder]
        ch.set_beta(beta)
        ch.run(2000)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        a.record()
        rounds = 0
        for k in range(iters // swap_every):
            ch.run_device(swap_every, 0, None)
            dev.sync()
            if mode == "ladders_local":
                ch.swap_step(T, rounds)
            else:
                P.tempered_swap(ch, c0, ladder, rounds)
            rounds += 1
        dev.sync()
        b.record()
        torch.cuda.synchronize()
        ms = tmax(a.elapsed_time(b), tdev)
        wall = tmax(time.perf_counter() - t0, tdev)
        cnts = ch.counters().sum(0)
        tot_c = torch.tensor([float(cnts[8]), float(cnts[9])], dtype=torch.float64, device=tdev)
        if world > 1:
            dist.all_reduce(tot_c)
        out["cfg4_" + mode] = {"chains": total, "per_gpu": nl, "ladder": ladder, "iters": iters, "swap_every": swap_every,
                               "wall_s": wall, "device_ms": ms, "it_per_s": total * iters / wall, "swap_accept": float(tot_c[1] / max(float(tot_c[0]), 1))}
        ch.close()
        if rank == 0:
            print(json.dumps({"cfg4_" + mode: out["cfg4_" + mode]}), file=sys.stderr, flush=True)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
