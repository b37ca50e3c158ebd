"""Multi-GPU execution on one 8xB200 box: one process per GPU, torch.distributed for the plumbing.

The path shards two ways:

* CHAINS are independent Markov chains: block-partition the chain ids over ranks, replicate the (tiny) binned
  statistics, no data-path collective.  A chain's Philox stream is keyed by (seed, global chain id), so results do
  not depend on the number of GPUs.
* LINEAGES, for very large tables: every rank bins a contiguous slice of (ts, te) into the raw integer accumulators
  of lr_bin_accumulate, one SUM all-reduce (NCCL over NVLink; int64, 8 x (n_bins+1) words per replicate -- 12.9 KB at
  200 bins) combines them, lr_bin_finalize reconstructs (sp, ex, br) on every rank.  All accumulators are integers
  (fractions of a year in 2^-52 fixed point), so the result is bit-identical for any number of ranks.
* TEMPERED ladders that span ranks exchange 16 bytes per chain (likelihood, inverse temperature) with one all-gather
  per swap round; states never cross NVLink, temperatures do.

The reference has no counterpart (its "multi-chain" mode is running the script several times by hand,
3_interpreting_literate_results_final.ipynb:92).
"""
from __future__ import annotations

import os

import numpy as np


def shard_range(n: int, world: int, rank: int):
    """Contiguous block partition of range(n): returns (start, count); the first n % world ranks hold one more."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, extra = divmod(int(n), world)
    start = rank * base + min(rank, extra)
    return start, base + (1 if rank < extra else 0)


def env_world():
    """(rank, local_rank, world_size) from the torchrun environment; (0, 0, 1) outside torchrun."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init(backend=None):
    """Join the process group torchrun describes (NCCL when CUDA is present, gloo otherwise)."""
    import torch
    import torch.distributed as dist
    rank, local, world = env_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, local, world


def allreduce_accumulators(acc, group=None):
    """SUM all-reduce of the raw K1 accumulators (int64 tensor, any shape), in place.  The only collective of the
    lineage-sharded path; issued once per dataset, never per iteration."""
    import torch
    import torch.distributed as dist
    if acc.dtype != torch.int64:
        raise TypeError("accumulators are int64")
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)
    return acc


def global_window(ts_local, te_local, group=None):
    """(first_bin, n_bins, start_time, end_time) of the WHOLE table from per-rank slices (torch tensors):
    range(int(min ts), int(max te)) of LiteRateForward.py:519 needs the global extrema."""
    import torch
    import torch.distributed as dist
    lo = ts_local.min().reshape(1).to(torch.float64)
    hi = te_local.max().reshape(1).to(torch.float64)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
    lo, hi = float(lo.item()), float(hi.item())
    return int(lo), int(hi) - int(lo), lo, hi


def bin_stats_lineage_sharded(dev, ts_local, te_local, first_bin, n_bins, fe_ref=0.5, dead_only=False, end_time=0.0,
                              stream=None, group=None):
    """Lineage-sharded K1: ts_local/te_local are this rank's float64 CUDA slices [n_local] or [n_rep, n_local].
    Returns (sp, ex, br) CUDA tensors [n_rep, n_bins], identical on every rank."""
    import torch
    if ts_local.dim() == 1:
        ts_local, te_local = ts_local[None, :], te_local[None, :]
    acc = dev.new_accumulators(ts_local.shape[0], n_bins, ts_local.device)
    if ts_local.shape[1] > 0:
        dev.bin_accumulate_device(ts_local, te_local, first_bin, n_bins, acc, fe_ref=fe_ref, dead_only=dead_only,
                                  end_time=end_time, stream=stream)
    if stream is not None:
        torch.cuda.current_stream().wait_stream(torch.cuda.ExternalStream(int(stream)))
    allreduce_accumulators(acc, group)
    if stream is not None:
        torch.cuda.ExternalStream(int(stream)).wait_stream(torch.cuda.current_stream())
    return dev.bin_finalize_device(acc, n_bins, fe_ref=fe_ref, stream=stream)


def gather_swap_info(info_local, group=None):
    """All-gather of the per-chain (likelihood, beta) pairs of a tempered ensemble: [n_local, 2] -> [world * n_local, 2]."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return info_local
    out = torch.empty((dist.get_world_size(group) * info_local.shape[0],) + tuple(info_local.shape[1:]),
                      dtype=info_local.dtype, device=info_local.device)
    dist.all_gather_into_tensor(out, info_local.contiguous(), group=group)
    return out


def tempered_swap(chains, first: int, ladder: int, round_: int, group=None):
    """One swap round of an ensemble whose ladders may span ranks: local (lik, beta) table -> all-gather (16 B per
    chain over NCCL) -> every rank applies the round to its own shard.  Every rank must hold the same number of chains."""
    import torch
    info = torch.empty((chains.n_chains, 2), dtype=torch.float64, device=torch.device("cuda", chains.dev.index))
    chains.swap_info_device(info)
    allinfo = gather_swap_info(info, group)
    chains.swap_apply_device(allinfo, first, ladder, round_)
    return allinfo


def temperature_ladder(n_temps: int, delta: float = 0.1):
    """Incremental-heating ladder beta_k = 1 / (1 + delta * k) (MrBayes/PyRate convention), beta_0 = 1 = the reference's chain."""
    return 1.0 / (1.0 + delta * np.arange(n_temps, dtype=np.float64))
