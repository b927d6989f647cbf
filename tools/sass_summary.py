"""Per-kernel counts of the Blackwell instructions that prove the tcgen05 / TMEM / bulk-copy path in the shipped library:
python tools/sass_summary.py [nerf-ca_b200/libnerfca_b200.so] > profiles/rXX/sass_summary.txt"""
import collections, re, subprocess, sys
so = sys.argv[1] if len(sys.argv) > 1 else "nerf-ca_b200/libnerfca_b200.so"
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
pats = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UBLKPF", "UTMALDG", "UTMASTG", "SYNCS", "MUFU.SIN", "MUFU.COS", "RED.E", "STL", "LDL", "USETMAXREG"]
cur, counts = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    for p in pats:
        if re.search(r"\b" + re.escape(p), line):
            counts[cur][p] += 1
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
        counts[cur]["instructions"] += 1
print(f"# {so}: cuobjdump -sass, instruction counts per kernel (UTCHMMA = tcgen05.mma kind::f16, LDTM/STTM = tcgen05.ld/st, UTCBAR = tcgen05.commit,")
print("# UBLKCP = cp.async.bulk (1-D TMA), UBLKPF = bulk L2 prefetch, SYNCS = mbarrier ops, STL/LDL = local-memory spills)")
print(f"{'kernel':60s} " + " ".join(f"{p:>9s}" for p in ["instructions"] + pats))
for k, c in counts.items():
    print(f"{k[:60]:60s} " + " ".join(f"{c[p]:9d}" for p in ["instructions"] + pats))
