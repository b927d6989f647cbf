"""Pretty-print a kernel timeline written by NERFCA_TIMELINE: python tools/tl_show.py log [first] [count]"""
import sys
lines = [l.split() for l in open(sys.argv[1]) if l.startswith('TL')]
ev = sorted(((int(a), int(b)) for _, a, b in lines), key=lambda e: e[1])
gaps = [i for i in range(1, len(ev)) if ev[i][1] - ev[i - 1][1] > 1000000]
ev = ev[gaps[-1]:] if gaps else ev
t0 = ev[0][1]
first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
count = int(sys.argv[3]) if len(sys.argv) > 3 else 80
prev = {}
for tag, t in ev[first:first + count]:
    who = {1: 'EPI ', 2: 'LOAD', 3: 'MMA '}.get(tag // 1000, '?   ')
    d = t - prev.get(tag // 1000, t)
    prev[tag // 1000] = t
    print(f"{t - t0:9d} {who} {tag:5d}  (+{d})")
print("events", len(ev), "span", ev[-1][1] - t0)
