"""Top stall instructions of one kernel in an .ncu-rep: python tools/ncu_hot.py rep kernel_name [N]"""
import csv, subprocess, sys, io
rep, kname = sys.argv[1], sys.argv[2]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kname], capture_output=True, text=True).stdout
lines = out.splitlines()
# several kernel instances may be concatenated; take the first block
start = [i for i, l in enumerate(lines) if l.startswith('"Address"')][0]
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith('"Kernel Name"')), len(lines))
rows = list(csv.reader(io.StringIO("\n".join(lines[start:end]))))
hdr = rows[0]
ia, isrc, ismp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
data = []
for k, r in enumerate(rows[1:]):
    try:
        data.append((int(r[ismp]), int(r[iex]), k, r[isrc].strip()))
    except ValueError:
        pass
tot = sum(d[0] for d in data)
totex = sum(d[1] for d in data)
print(f"total samples {tot}, total warp-instructions executed {totex}")
for s, ex, k, src in sorted(data, reverse=True)[:N]:
    print(f"{100.0 * s / max(tot,1):6.2f}%  ex={ex:9d}  #{k:5d}  {src[:110]}")
