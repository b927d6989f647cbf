"""Executed warp-instructions of one kernel in an .ncu-rep, cumulated between SASS landmarks: python tools/ncu_regions.py rep kernel"""
import csv, subprocess, sys, io, re
rep, kname = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kname], capture_output=True, text=True).stdout
lines = out.splitlines()
start = [i for i, l in enumerate(lines) if l.startswith('"Address"')][0]
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith('"Kernel Name"')), len(lines))
rows = list(csv.reader(io.StringIO("\n".join(lines[start:end]))))
hdr = rows[0]
isrc, iex, ismp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
pat = re.compile(r"LDTM|STTM|UTCHMMA|UTCBAR|SYNCS|BAR\.SYNC|MUFU|DMUL|STG|EXIT|WARPSYNC")
tot = 0; acc = 0; accs = 0; n = 0
data = []
for k, r in enumerate(rows[1:]):
    try:
        ex, smp = int(r[iex]), int(r[ismp])
    except ValueError:
        continue
    data.append((k, r[isrc].strip(), ex, smp))
tot = sum(d[2] for d in data); tots = sum(d[3] for d in data)
print(f"total executed {tot}, samples {tots}")
for k, src, ex, smp in data:
    acc += ex; accs += smp; n += 1
    if pat.search(src):
        print(f"#{k:5d} +{n:4d} instrs  ex {acc:10d} ({100.0*acc/tot:5.1f}%)  samples {100.0*accs/max(tots,1):5.1f}%   | {src[:70]}")
        acc = 0; accs = 0; n = 0
print(f"tail +{n} ex {acc}")
