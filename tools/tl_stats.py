"""Per-interval statistics of a kernel timeline written by NERFCA_TIMELINE: python tools/tl_stats.py log top|bot"""
import bisect
import statistics as st
import sys


def load(path):
    lines = [l.split() for l in open(path) if l.startswith('TL')]
    ev = sorted(((int(a), int(b)) for _, a, b in lines), key=lambda e: e[1])
    gaps = [i for i in range(1, len(ev)) if ev[i][1] - ev[i - 1][1] > 1000000]
    return ev[gaps[-1]:] if gaps else ev


ev = load(sys.argv[1])
t = {}
for tag, c in ev:
    t.setdefault(tag, []).append(c)


def gap(a_tag, b_tag):
    r = []
    for a in t.get(a_tag, []):
        i = bisect.bisect_right(t.get(b_tag, []), a)
        if i < len(t.get(b_tag, [])):
            r.append(t[b_tag][i] - a)
    return r


role = sys.argv[2]
start = 3010 if role == 'top' else 3000
p = [b - a for a, b in zip(t[start], t[start][1:])]
print(f"{sys.argv[1]}: tiles {len(t[start])} period mean {st.mean(p):.0f} median {st.median(p):.0f} span {ev[-1][1] - ev[0][1]}")
pairs = {'top': [(3010, 1020), (1020, 1021), (1021, 1023), (1023, 1024), (3020, 1030), (1030, 1033), (1033, 1009), (1009, 1010), (1010, 1013), (1013, 1014), (2001, 2003), (2003, 2005)],
         'bot': [(3000, 3001), (3000, 1010), (1010, 1011), (1011, 1014), (3010, 3011), (3010, 1020), (1020, 1021), (1021, 1024), (3020, 3021), (3021, 3000), (2000, 2001), (2001, 2002), (2002, 2003), (2002, 3000)]}[role]
for a, b in pairs:
    g = gap(a, b)
    if g:
        print(f"  {a}->{b}: mean {st.mean(g):.0f} median {st.median(g):.0f} max {max(g)}")
