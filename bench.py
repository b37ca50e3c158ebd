#!/usr/bin/env python
"""bench.py -- RJMCMC iterations/s (all chains) at 1M lineages, and the K1 HBM roofline.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # the CUDA arm (one JSON line on rank 0)
    python bench.py --impl reference [--steps K] [--warmup W]      # the CPU arm (oracle port, all host cores)

Workload (BASELINE.json configs[2]): synthetic 1M lineages x 200 one-year bins, 256 chains on one B200,
-model_BDI 0, every chain on its own stochastic-imputation replicate of the table (the reference's published
workflow is "100 chains on 100 imputations"), i.e. 256 x 1M lineages = 4.1 GB of (ts, te) per step -- larger
than the 126 MB L2, so no flush is needed between steps.  One STEP = one pass of the hot path over that batch:

    K1  lineages -> per-bin (births, deaths, time at risk) for all replicates        lr_bin_accumulate + lr_bin_finalize
    K2  prefix tables of the likelihood                                               lr_dataset_create
    K3  ITERS iterations of every chain, one sample record every SAMPLE iterations    lr_chains_create + lr_chains_run

value = chains x ITERS / step time with (ts, te) resident in HBM; e2e = the same through the host-buffer API
(pinned host ts/te in, sample records out) with the copies inside the timed region.  With N GPUs every rank
runs its own 256 chains on its own replicates (weak scaling, no data-path collective).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

N_LINEAGES = 1_000_000
N_BINS = 200
FIRST_BIN = 1800
CHAINS = 256
ITERS = 100_000          # iterations of every chain per step
SAMPLE = 1000            # the reference's default -s
METRIC = "RJMCMC iters/sec (all chains) at 1M lineages"
UNIT = "it/s"


def _args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=10)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    p.add_argument("--chains", type=int, default=CHAINS, help="chains per GPU (default: the configuration the metric is quoted on)")
    p.add_argument("--lineages", type=int, default=N_LINEAGES)
    p.add_argument("--iters", type=int, default=ITERS)
    p.add_argument("--shared-dataset", type=int, default=0, help="1: all chains share one replicate (n_rep = 1)")
    p.add_argument("--real", type=int, default=0, help="1: real-valued times (syn-real, -death_jitter 0) instead of integer years")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-e2e", action="store_true")
    return p.parse_args()


def _workload(a, n_gpus):
    n_rep = 1 if a.shared_dataset else a.chains
    return {
        "workload": "syn-%s %d lineages x %d one-year bins, %d chains/GPU, %s, model_BDI 0, %d it/chain/step, sample every %d"
                    % ("real" if a.real else "int", a.lineages, N_BINS, a.chains,
                       "one imputation replicate per chain" if n_rep > 1 else "one shared table", a.iters, SAMPLE),
        "lineages": a.lineages, "bins": N_BINS, "chains_per_gpu": a.chains, "replicates_per_gpu": n_rep,
        "iters_per_chain_per_step": a.iters, "sample_every": SAMPLE,
        "parallelism": "chains sharded over %d GPU(s), no collective" % n_gpus,
        "l2": "inputs (%.2f GB per step) larger than L2, no flush" % (16.0 * a.lineages * n_rep / 1e9)
              if 16.0 * a.lineages * n_rep > 256e6 else "L2 flushed (512 MB memset) before every step",
    }


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.t = threading.Thread(target=self._read, daemon=True)
        self.t.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        rows = [r for (t, r) in self.rows if t0 <= t <= t1 and len(r) >= 8] or [r for (_, r) in self.rows if len(r) >= 8]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = [float(r[1]) for r in rows]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[4 + k].lower().startswith("active") for r in rows)]
        pw = [float(r[3]) for r in rows if r[3].replace(".", "", 1).isdigit()]
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": float(rows[0][2]), "reasons": reasons,
                "samples": len(rows), "power_w_max": max(pw) if pw else None}


def _peaks():
    try:
        with open(os.path.join(REPO, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (copy kernel, read+write)"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s of B200_PROFILING.md (MEASURED_PEAKS.json absent)"


def _traffic(a):
    """dram bytes per K1 launch from the committed ncu --set full capture of this workload, if there is one."""
    try:
        with open(os.path.join(REPO, "profiles", "k1_traffic.json")) as fh:
            t = json.load(fh)
        key = "%s_%d_x_%d" % ("real" if a.real else "int", a.lineages, 1 if a.shared_dataset else a.chains)
        return t.get(key)
    except Exception:
        return None


# ------------------------------------------------------------------------------------------ CPU arm
def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import cpu_baseline as CB
    cores = os.cpu_count() or 1
    n_rep = 1 if a.shared_dataset else a.chains
    bins, iters = 2, 4000
    for _ in range(a.warmup):
        CB.run_sample(cores, a.lineages, N_BINS, n_rep, a.chains, a.iters, SAMPLE, bins=1, iters=500)
    t0 = time.perf_counter()
    res = [CB.run_sample(cores, a.lineages, N_BINS, n_rep, a.chains, a.iters, SAMPLE, bins=bins, iters=iters, seed0=1 + 97 * k)
           for k in range(a.steps)]
    wall = time.perf_counter() - t0
    full = statistics.mean(r["whole_workload_s"] for r in res)
    value = a.chains * a.iters / full
    sample = ("per step every one of %d worker processes bins %d of the %d unit bins of one %d-lineage replicate with the reference's "
              "per-bin NumPy formulation and runs %d RJMCMC iterations of one chain (scipy priors, adequacy on, sampling every %d); "
              "extrapolated linearly to %d replicates x %d bins and %d chains x %d iterations" %
              (cores, bins, N_BINS, a.lineages, iters, SAMPLE, n_rep, N_BINS, a.chains, a.iters))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": 1e3 * wall / max(a.steps, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": _workload(a, a.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "loop_it_per_s_1core": statistics.mean(r["it_per_s_loop_1core"] for r in res),
                         "binning_s_per_bin_1core": statistics.mean(r["s_per_bin_1core"] for r in res),
                         "whole_workload_s": full},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------ CUDA arm
def run_cuda(a):
    import numpy as np
    import torch
    import torch.distributed as dist
    from literate_b200 import engine as E, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; literate_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    tdev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=tdev)
    n_gpus = world

    dev = E.Device(local)
    n, nb, chains = a.lineages, N_BINS, a.chains
    n_rep = 1 if a.shared_dataset else chains
    fe_ref = 1.0 if a.real else 0.5
    end_time = float(FIRST_BIN + nb) + (0.0 if a.real else 0.5)
    rep0 = rank * n_rep                                    # every rank has its own replicates and chain ids
    gen = synth.syn_real_device if a.real else synth.syn_int_device
    ts, te = gen(n, n_rep, tdev, seed=synth.BASE_SEED + 1000 * rank)
    ts, te = ts[:, :n], te[:, :n]
    rep_of_chain = (np.arange(chains) % n_rep).astype(np.int32)
    n_rec = (a.iters + SAMPLE - 1) // SAMPLE
    records = torch.empty((n_rec, chains, E.LR_REC_DOUBLES), dtype=torch.float64, device=tdev)
    acc = dev.new_accumulators(n_rep, nb, tdev)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=tdev) if 16.0 * n * n_rep <= 256e6 else None
    stream = torch.cuda.Stream(device=tdev)
    cfg = E.default_config(0)
    k1_ms, k3_ms = [], []

    def step(k, timed):
        """One pass of the hot path with device-resident lineages."""
        if flush is not None:
            flush.zero_()
        acc.zero_()
        e0, e1, e2, e3 = (torch.cuda.Event(enable_timing=True) for _ in range(4))
        e0.record()
        dev.bin_accumulate_device(ts, te, FIRST_BIN, nb, acc, fe_ref=fe_ref, stream=stream.cuda_stream)
        e1.record()
        sp, ex, br = dev.bin_finalize_device(acc, nb, fe_ref=fe_ref, stream=stream.cuda_stream)
        ds = E.Dataset.from_device(dev, sp, ex, br, 0, float(FIRST_BIN), end_time, stream=stream.cuda_stream)
        ch = E.Chains(ds, chains, seed=2026 + k, cfg=cfg, chain_id0=rank * chains, rep_of_chain=rep_of_chain)
        e2.record()
        ch.run_device(a.iters, SAMPLE, records, stream=stream.cuda_stream)
        e3.record()
        return (e0, e1, e2, e3), ds, ch

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.cuda.stream(stream):
        for k in range(a.warmup):
            _, ds, ch = step(k, False)
            torch.cuda.synchronize()
            ch.close(); ds.close()
        barrier()
        clocks = ClockSampler(local)
        clocks.start()
        launches0 = dev.kernel_launches
        t_wall0 = time.perf_counter()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        keep = []
        for k in range(a.steps):
            keep.append(step(a.warmup + k, True))
        g1.record()
        barrier()
        t_wall1 = time.perf_counter()
        total_ms = g0.elapsed_time(g1)
        launches = dev.kernel_launches - launches0
        lik_evals = 0
        for (e0, e1, e2, e3), ds, ch in keep:
            k1_ms.append(e0.elapsed_time(e1))
            k3_ms.append(e2.elapsed_time(e3))
            lik_evals += int(ch.counters()[:, 2].sum())
            ch.close(); ds.close()
        # sanity of the last step's output: every chain delivered its samples, counts conserve lineages
        rec = records.cpu().numpy()
        assert np.all(rec[:, :, E.REC_IT] == (np.arange(n_rec) * SAMPLE)[:, None]) and np.all(np.isfinite(rec[:, :, E.REC_LIK]))

        # ---------------- e2e: host buffers through the public API, copies inside the timed region
        e2e = None
        if not a.no_e2e:
            hts = torch.empty((n_rep, n), dtype=torch.float64, pin_memory=True); hts.copy_(ts)
            hte = torch.empty((n_rep, n), dtype=torch.float64, pin_memory=True); hte.copy_(te)
            hrec = torch.empty((n_rec, chains, E.LR_REC_DOUBLES), dtype=torch.float64, pin_memory=True)
            nts, nte, nrec = hts.numpy(), hte.numpy(), hrec.numpy()

            # the public streaming API: engine.Pipeline double-buffers consecutive batches (tables of step k+1 are copied and
            # binned while the chains of step k run); every step's H2D copy and D2H read are inside the timed region,
            # which ends when the last step's records are on the host
            pipe = E.Pipeline(local)

            def e2e_push(k):
                return pipe.push(nts, nte, chains, a.iters, SAMPLE, seed=4052 + k, cfg=cfg, first_bin=FIRST_BIN, n_bins=nb,
                                 death_jitter=0.0 if a.real else 0.5, start_time=float(FIRST_BIN), end_time=end_time,
                                 rep_of_chain=rep_of_chain, chain_id0=rank * chains, out=hrec)
            for k in range(max(1, min(a.warmup, 2))):
                e2e_push(k)
            pipe.flush(out=hrec)
            barrier()
            n_e2e = max(2, a.steps)                  # as many steps as the device-timed value; the last step's chains cannot overlap a copy
            t0 = time.perf_counter()
            for k in range(n_e2e):
                tp = time.perf_counter()
                e2e_push(100 + k)
                if os.environ.get("LR_BENCH_DEBUG"):
                    print("push %d: %.1f ms" % (k, 1e3 * (time.perf_counter() - tp)), file=sys.stderr, flush=True)
            tp = time.perf_counter()
            pipe.flush(out=hrec)
            tq = time.perf_counter()
            torch.cuda.synchronize()
            e2e_s = time.perf_counter() - t0
            if os.environ.get("LR_BENCH_DEBUG"):
                print("flush %.1f ms, sync %.1f ms, total %.1f ms" % (1e3 * (tq - tp), 1e3 * (time.perf_counter() - tq), 1e3 * e2e_s), file=sys.stderr, flush=True)
            assert np.all(nrec[:, :, E.REC_IT] == (np.arange(n_rec) * SAMPLE)[:, None]) and np.all(np.isfinite(nrec[:, :, E.REC_LIK]))
            pipe.close()
            e2e = [e2e_s / n_e2e, 16 * n * n_rep + 4 * chains, int(nrec.nbytes) + 24 * n_rep * nb]
            del hts, hte, hrec
        cl = clocks.stop(t_wall0, t_wall1)

        # ---------------- the sibling samplers of SURVEY 8 f-4 on the statistics of replicate 0 (reported beside the metric, not part of it)
        siblings = None
        if rank == 0:
            from literate_b200 import trend as TR, ddrate as DD
            sp, ex, br = (x[0].cpu().numpy() for x in dev.bin_finalize_device(acc, nb, fe_ref=fe_ref, stream=stream.cuda_stream))
            trend = np.clip(np.linspace(0.0, 1.0, nb), 1e-15, 1.0)
            sib_iters = 20000
            siblings = {"chains": chains, "bins": nb, "iterations": sib_iters}
            for name, mk in (("k6_trend_it_per_s", lambda: TR.TrendChains(dev, sp, ex, br, trend, chains, 7)),
                             ("k7_ddrate_it_per_s", lambda: DD.DDChains(dev, sp, ex, br, float(FIRST_BIN), end_time, 2, 2, None, None, chains, 7))):
                sc = mk()
                sc.run(2000)
                srec = torch.empty((sc.records_per_run(sib_iters, SAMPLE), chains, sc.rec_doubles), dtype=torch.float64, device=tdev)
                s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s0.record(); sc.run_device(sib_iters, SAMPLE, srec, stream=stream.cuda_stream); s1.record()
                torch.cuda.synchronize()
                siblings[name] = chains * sib_iters / (s0.elapsed_time(s1) * 1e-3)
                assert bool(torch.isfinite(srec[:, :, 1]).all())
                sc.close()

    t = torch.tensor([total_ms, e2e[0] if e2e else 0.0], dtype=torch.float64, device=tdev)
    tot = torch.tensor([float(lik_evals), float(launches)], dtype=torch.float64, device=tdev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    total_ms, e2e_step_s = float(t[0]), float(t[1])
    if rank == 0:
        peak, peak_src = _peaks()
        k1 = statistics.mean(k1_ms)
        algo_bytes = 16.0 * n * n_rep
        achieved = algo_bytes / (k1 * 1e-3) / 1e9
        value = n_gpus * chains * a.iters * a.steps / (total_ms * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": total_ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": _workload(a, n_gpus), "clocks": cl,
            "gpu_launches": int(tot[1]),
            "lik_evals_per_s": float(tot[0]) / (total_ms * 1e-3),
            "kernels": {"k1_bin_kernel_ms": k1, "k3_run_kernel_ms": statistics.mean(k3_ms),
                        "k3_ns_per_iteration_per_chain": 1e6 * statistics.mean(k3_ms) / a.iters,
                        "k3_bound": "dependent-instruction latency of one chain warp per chain (no DRAM traffic in the loop; see profiles/)",
                        "k1_share_of_step": k1 * a.steps / total_ms, "k3_share_of_step": sum(k3_ms) / total_ms},
            "roofline": {"kernel": "k1_bin_kernel (lineages -> per-bin births/deaths/time at risk)", "bound": "hbm", "achieved": achieved,
                         "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": _traffic(a),
                         "algorithmic_bytes_per_launch": algo_bytes, "peak_source": peak_src, "ms_per_launch": k1},
        }
        if siblings:
            line["siblings"] = siblings
        if e2e:
            line["e2e"] = {"value": n_gpus * chains * a.iters / e2e_step_s, "unit": UNIT, "h2d_bytes_per_step": e2e[1],
                           "d2h_bytes_per_step": e2e[2], "ms_per_step": 1e3 * e2e_step_s}
        if n_gpus == 1 and not a.no_cpu_baseline:
            from oracle import cpu_baseline as CB
            cores = os.cpu_count() or 1
            bins, iters = 2, 4000
            r = CB.run_sample(cores, n, nb, n_rep, chains, a.iters, SAMPLE, bins=bins, iters=iters)
            line["cpu_baseline"] = {
                "value": r["value"], "unit": UNIT, "cores": cores, "kind": "port",
                "sample": "one pass: each of %d worker processes bins %d of %d bins of one %d-lineage replicate (reference's per-bin NumPy "
                          "formulation) and runs %d iterations of one chain (scipy priors); extrapolated linearly to the whole workload"
                          % (cores, bins, nb, n, iters),
                "loop_it_per_s_1core": r["it_per_s_loop_1core"], "binning_s_per_bin_1core": r["s_per_bin_1core"],
                "whole_workload_s": r["whole_workload_s"]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    args = _args()
    # stdout carries exactly one JSON line: native libraries that print there (NCCL's version banner) go to stderr
    sys.stdout.flush()
    _json_fd = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(_json_fd, "w")
    sys.exit(run_reference(args) if args.impl == "reference" else run_cuda(args))
