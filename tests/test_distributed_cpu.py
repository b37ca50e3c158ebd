"""CPU, world_size 2 over gloo: the host-side logic of the multi-GPU paths.

The kernels cannot run here, so the per-rank raw accumulators come from the oracle's integer restatement of
lr_bin_accumulate; what is under test is the product's sharding arithmetic and collectives (literate_b200.parallel):
a lineage-sharded table, all-reduced, finalises to exactly the statistics of the whole table.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import REPO
from oracle import literate_oracle as O
from literate_b200 import parallel as P, synth


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 256, 4096, 1_000_003):
        for world in (1, 2, 3, 8):
            parts = [P.shard_range(n, world, r) for r in range(world)]
            assert parts[0][0] == 0 and sum(c for _, c in parts) == n
            for (s0, c0), (s1, _) in zip(parts, parts[1:]):
                assert s0 + c0 == s1
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1
    with pytest.raises(ValueError):
        P.shard_range(5, 2, 2)


def test_temperature_ladder():
    b = P.temperature_ladder(4, 0.1)
    assert b[0] == 1.0 and np.all(np.diff(b) < 0) and b[3] == pytest.approx(1 / 1.3)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _rank_main(rank, world, port, real, q):
    sys.path.insert(0, REPO)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    r, _, w = P.init("gloo")
    assert (r, w) == (rank, world)
    n = 6001
    ts, te = synth.syn_real(n, replicate=3) if real else synth.syn_int(n, replicate=3)
    fe_ref = 1.0 if real else 0.5
    s0, cnt = P.shard_range(n, world, rank)
    lts, lte = torch.from_numpy(ts[s0:s0 + cnt]), torch.from_numpy(te[s0:s0 + cnt])
    first, nb, lo, hi = P.global_window(lts, lte)
    acc = torch.from_numpy(O.raw_accumulators(ts[s0:s0 + cnt], te[s0:s0 + cnt], first, nb, fe_ref))
    P.allreduce_accumulators(acc)
    sp, ex, br = O.finalize_accumulators(acc.numpy(), nb, fe_ref)
    # tempered-swap exchange: every rank ends up with everyone's (lik, beta) rows in rank order
    info = torch.full((4, 2), float(rank), dtype=torch.float64)
    allinfo = P.gather_swap_info(info)
    q.put((rank, first, nb, lo, hi, sp, ex, br, allinfo.numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("real", [0, 1])
def test_lineage_sharded_binning_world2(real):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_rank_main, args=(r, world, port, real, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    ts, te = synth.syn_real(6001, replicate=3) if real else synth.syn_int(6001, replicate=3)
    want = O.bin_stats(ts, te)
    for rank, first, nb, lo, hi, sp, ex, br, allinfo in outs:
        assert (first, nb) == (want.first_bin, want.n_bins) and lo == ts.min() and hi == te.max()
        assert (sp == want.sp).all() and (ex == want.ex).all()
        if real:
            np.testing.assert_allclose(br, want.br, rtol=1e-12)
        else:
            assert (br == want.br).all()
        assert allinfo.shape == (8, 2) and (allinfo[:4] == 0).all() and (allinfo[4:] == 1).all()
    assert (outs[0][7] == outs[1][7]).all()          # bit-identical on every rank, also for real-valued times
