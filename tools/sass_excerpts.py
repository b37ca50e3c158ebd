"""profiles/sass/: SASS of the hot kernels of the BUILT library (cuobjdump -sass), so that the claims in profiles/*.md can be
checked from the repository alone: full listings of the two K1 kernels, and for the (large) K3 kernels the opcode histogram,
the register / stack line of ptxas and the lines that show the mechanisms the text names (ATOMS.POPC.INC, LDG...NA..128,
REDUX, MATCH, NANOSLEEP, LD/ST...STRONG/volatile shared accesses of the team protocol, USETMAXREG).  Run after build()."""
import collections, hashlib, os, re, subprocess, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, "literate_b200", "_lib", "libliterate_b200.so")
OUT = os.path.join(REPO, "profiles", "sass")
os.makedirs(OUT, exist_ok=True)
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
funcs, cur = {}, None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); funcs[cur] = []
    elif cur is not None:
        funcs[cur].append(line)
digest = hashlib.sha256(open(LIB, "rb").read()).hexdigest()[:16]
stamp = open(os.path.join(REPO, "literate_b200", "_lib", "build.sha256")).read().strip()[:16]
log = open(os.path.join(REPO, "literate_b200", "_lib", "build.log")).read()


def ptxas_line(name):
    m = re.search(re.escape(name) + r"[^\n]*\n[^\n]*\n(ptxas info\s+: Used[^\n]*)", log)
    return m.group(1) if m else ""


def instrs(lines):
    out = []
    for l in lines:
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            out.append((m.group(1), m.group(2).strip()))
    return out


index = ["library sha256[:16] %s, source digest stamp %s" % (digest, stamp), ""]
for name, lines in funcs.items():
    short = None
    for key in ("k1_bin_kernel", "k1_bin_i32_kernel", "k1_bin_lanes_kernelILb0ELi4", "k3_team_kernelILi8", "k3_team_kernelILi16", "k3_team_kernelILi4", "k3_run_kernelILi1", "k3_run_kernelILi0"):
        if key in name:
            short = key
    if short is None:
        continue
    ins = instrs(lines)
    hist = collections.Counter((i.split()[1] if i.startswith("@") else i.split()[0]).split(".")[0] for _, i in ins)
    index.append("%s: %d instructions; %s" % (short, len(ins), ptxas_line(name)))
    with open(os.path.join(OUT, short + ".txt"), "w") as fh:
        fh.write("# %s\n# %s\n# %s\n" % (name, ptxas_line(name), index[0]))
        fh.write("# opcode histogram (static): " + ", ".join("%s %d" % kv for kv in hist.most_common(25)) + "\n")
        if short.startswith("k1_"):
            fh.write("\n".join("%s  %s" % a for a in ins) + "\n")
        else:
            pat = re.compile(r"ATOMS|ATOMG|RED\.|REDUX|MATCH|NANOSLEEP|USETMAXREG|LDS.*64|STS.*64|LD\.E.*STRONG|ST\.E.*STRONG|MEMBAR|BAR\.|LDL|STL|WARPSYNC|CREDUX")
            fh.write("# lines that show the mechanisms named in profiles/r02_k3_team_and_pipeline.md (address, instruction):\n")
            fh.write("\n".join("%s  %s" % a for a in ins if pat.search(a[1])) + "\n")
with open(os.path.join(OUT, "INDEX.txt"), "w") as fh:
    fh.write("\n".join(index) + "\n")
print("\n".join(index))
