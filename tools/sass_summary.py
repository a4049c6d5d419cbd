"""Per-kernel SASS mnemonic summary of libprismer_sm100.so (what proves a Blackwell-native kernel: UTC*MMA = tcgen05.mma, LDTM/STTM =
tcgen05.ld/st, UTMALDG/UTMASTG = TMA, HMMA = mma.sync, LDGSTS = cp.async).      python tools/sass_summary.py > profiles/sass_r2_summary.txt"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "prismer_b200", "libprismer_sm100.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "HMMA", "LDGSTS", "LDSM", "MUFU.EX2", "SYNCS", "STG.E.128", "LDG.E.128",
        "RED", "ATOM"]
cur, stats = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        stats[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        stats[cur]["_n"] += 1
        for k in KEYS:
            if op.startswith(k):
                stats[cur][k] += 1
demangle = subprocess.run(["c++filt"], input="\n".join(stats), capture_output=True, text=True).stdout.splitlines()
print(f"# SASS mnemonic counts per kernel of {os.path.relpath(lib, ROOT)} (cuobjdump -sass; sm_100a)")
print(f"# {'kernel':78s} {'instr':>6s}  " + "  ".join(f"{k}" for k in KEYS))
for (name, c), dm in zip(stats.items(), demangle):
    short = re.sub(r"\(anonymous namespace\)::", "", dm)
    short = re.sub(r"\(.*", "", short)[:78]
    print(f"{short:80s} {c['_n']:6d}  " + "  ".join(f"{c[k]:{len(k)}d}" for k in KEYS))
