"""Host -> device copy bandwidth with all ranks copying at once: is the end-to-end step of bench.py at the machine's limit?

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/h2d_probe.py

Every rank copies a pinned 1 GiB buffer to its GPU 20 times, all ranks between the same two barriers; one cudaMemcpyAsync per copy.
Two passes: (a) buffers allocated as the process starts; (b) the process first pins itself to the CPUs NVML reports as local to
its GPU (nvmlDeviceGetCpuAffinity), so that first touch places the pinned pages on the GPU's NUMA node.  Rank 0 prints one JSON
object with per-rank and aggregate GB/s and the topology facts it could read (NUMA node of every GPU, CPUs per node)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist


def gpu_cpus(index):
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        n = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n)
        cpus = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1]
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        node = None
        p = "/sys/bus/pci/devices/%s/numa_node" % bus.lower()[-12:]
        if os.path.exists(p):
            node = int(open(p).read())
        return cpus, node, bus
    except Exception as e:                      # noqa: BLE001
        return None, None, "nvml unavailable: %s" % e


def measure(tdev, rank, world, nbytes=1 << 30, reps=20):
    host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    host.fill_(rank + 1)                        # first touch by this process
    devb = torch.empty(nbytes, dtype=torch.uint8, device=tdev)
    for _ in range(3):
        devb.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    for _ in range(reps):
        devb.copy_(host, non_blocking=True)
    b.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    mine = nbytes * reps / (a.elapsed_time(b) * 1e-3) / 1e9
    t = torch.tensor([mine, wall], dtype=torch.float64, device=tdev)
    allr = [torch.zeros_like(t) for _ in range(world)]
    if world > 1:
        dist.all_gather(allr, t)
    else:
        allr = [t]
    rates = [float(x[0]) for x in allr]
    slowest = max(float(x[1]) for x in allr)
    del host, devb
    return {"per_rank_GBps": rates, "aggregate_GBps": world * nbytes * reps / slowest / 1e9}


def main():
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    tdev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=tdev)
    out = {"world": world, "cpus": os.cpu_count()}
    cpus, node, bus = gpu_cpus(local)
    facts = [cpus and len(cpus), node, bus]
    out["default_affinity"] = measure(tdev, rank, world)
    if cpus:
        try:
            os.sched_setaffinity(0, cpus)
            out["gpu_local_affinity"] = measure(tdev, rank, world)
        except OSError as e:
            out["gpu_local_affinity"] = "sched_setaffinity failed: %s" % e
    gathered = [None] * world
    if world > 1:
        dist.all_gather_object(gathered, facts)
    else:
        gathered = [facts]
    if rank == 0:
        out["gpu_local_cpus_numa_node_bus"] = gathered
        try:
            out["numa_nodes"] = sorted(d for d in os.listdir("/sys/devices/system/node") if d.startswith("node"))
        except OSError:
            out["numa_nodes"] = None
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
