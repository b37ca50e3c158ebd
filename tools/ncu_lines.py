"""Per-source-line instruction counts and stall samples from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`.
python tools/ncu_lines.py report.csv [top N] [divide executed instructions by this number, e.g. chains x iterations]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
div = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
fpath, hdr = None, None
agg = collections.OrderedDict()
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fpath = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r; continue
    if hdr is None:
        continue
    if r[0] != "":
        key = (fpath, int(r[0]), r[1].strip()[:110])
        agg.setdefault(key, [0.0, 0.0])
        cur = key
        continue
    ix = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples")
    try:
        agg[cur][0] += float(r[ix] or 0); agg[cur][1] += float(r[isamp] or 0)
    except (ValueError, IndexError):
        pass
tot_i = sum(v[0] for v in agg.values()); tot_s = sum(v[1] for v in agg.values())
print("total warp instructions %.4g (%.1f per unit), samples %d" % (tot_i, tot_i / div, tot_s))
byfile = collections.Counter()
for (f, l, s), v in agg.items():
    byfile[f] += v[0]
print({k: round(v / div, 1) for k, v in byfile.items()})
for (f, l, s), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%7.1f instr %5.1f%% | samples %5.1f%% | %s:%d  %s" % (v[0] / div, 100 * v[0] / tot_i, 100 * v[1] / max(tot_s, 1), f, l, s))
