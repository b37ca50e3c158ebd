"""Turn one ncu capture of `python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e` into the two small
JSON files bench.py reads (with the digest of the kernel sources, so that a stale capture is never reported):

    ncu --set full --clock-control none -k regex:"k1_bin|k3_" -o gpurun_out/r02_bench python bench.py ...          (GPU box)
    ncu -i gpurun_out/r02_bench.ncu-rep --page raw --csv > profiles/r02_bench_raw.csv                              (here)
    python tools/profile_to_json.py profiles/r02_bench_raw.csv [chains] [iterations per step] [steps captured]

profiles/k1_traffic.json       dram__bytes_read.sum + dram__bytes_write.sum per K1 launch
profiles/k3_instructions.json  smsp__inst_executed.sum of all K3 kernels of one step / (chains x iterations)
"""
import csv, hashlib, json, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sha16(rel):
    with open(os.path.join(REPO, rel), "rb") as fh:
        return hashlib.sha256(fh.read()).hexdigest()[:16]


rows = list(csv.reader(open(sys.argv[1])))
chains = int(sys.argv[2]) if len(sys.argv) > 2 else 256
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 100000
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
name_i = col["Kernel Name"]
k1, k1_lanes, k3_inst, k3_time = [], [], 0.0, 0.0


def dram_bytes(r):
    def unit(c):
        return {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[rows[1][col[c]]]
    return float(r[col["dram__bytes_read.sum"]]) * unit("dram__bytes_read.sum") + float(r[col["dram__bytes_write.sum"]]) * unit("dram__bytes_write.sum")


k3_steps = 0
for r in rows[2:]:
    if len(r) <= name_i:
        continue
    nm = r[name_i]
    if "k1_bin_lanes_kernel" in nm or "k1_bin_kernel" in nm:
        k1.append(dram_bytes(r))          # in launch order: the steps' passes, then bench.py's real-valued section (7 shuffled, 7 sorted)
    elif "k3_team_kernel" in nm or "k3_run_kernel" in nm:
        k3_inst += float(r[col["smsp__inst_executed.sum"]])
        k3_time += float(r[col["gpu__time_duration.sum"]])
        if "k3_team_kernel" in nm:
            k3_steps += 1
steps = int(sys.argv[4]) if len(sys.argv) > 4 else max(1, k3_steps // 2)       # two team passes per step
src = ("k3_chains.cu", "k3_team.cuh", "chain_device.cuh", "lr_common.cuh")
if k1:
    p = os.path.join(REPO, "profiles", "k1_traffic.json")
    t = json.load(open(p)) if os.path.exists(p) else {}
    t["_comment"] = "dram__bytes_read.sum + dram__bytes_write.sum per K1 launch (k1_bin_lanes_kernel at the bench sizes) from an ncu --set full capture of bench.py; " \
                    "key <table kind>_<lineages>_x_<replicates>; written by tools/profile_to_json.py with the digest of the kernel source"
    t["k1_binstats_cu_sha16"] = sha16("literate_b200/csrc/k1_binstats.cu")
    for key in ("real_lanes_1000000_x_%d" % chains, "realsorted_lanes_1000000_x_%d" % chains):
        t.pop(key, None)
    if len(k1) > 14 and "--no-real" not in sys.argv:
        k1, k1_lanes = k1[:-14], k1[-14:]
        t["real_lanes_1000000_x_%d" % chains] = int(sum(k1_lanes[:7]) / 7)
        t["realsorted_lanes_1000000_x_%d" % chains] = int(sum(k1_lanes[7:]) / 7)
    t["int_1000000_x_%d" % chains] = int(sum(k1) / len(k1))
    json.dump(t, open(p, "w"), indent=1)
    print("k1 traffic per launch:", int(sum(k1) / len(k1)), "from", len(k1), "launches; lane-private build:", len(k1_lanes), "launches")
if k3_inst:
    ipi = k3_inst / (steps * chains * iters)
    json.dump({"_comment": "warp instructions (smsp__inst_executed.sum) of all K3 kernels of one bench step / (chains x iterations), "
                           "ncu capture of bench.py; written by tools/profile_to_json.py",
               "warp_instructions_per_iteration": ipi, "chains": chains, "iterations_per_step": iters, "steps_captured": steps,
               "k3_ms_per_step_under_ncu": k3_time / steps * (1e-6 if rows[1][col["gpu__time_duration.sum"]] == "ns" else 1e-3 if rows[1][col["gpu__time_duration.sum"]] in ("us", "usecond") else 1),
               "sources_sha16": "+".join(sha16("literate_b200/csrc/" + f) for f in src)},
              open(os.path.join(REPO, "profiles", "k3_instructions.json"), "w"), indent=1)
    print("k3 warp instructions per iteration: %.1f over %d step(s)" % (ipi, steps))
