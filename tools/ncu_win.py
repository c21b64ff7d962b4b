"""Windowed view of an `ncu --page source --csv --print-source sass` export: share of stall samples per window of SASS
instructions with the landmark opcodes in it, then the hottest instructions.  usage: ncu_win.py file.csv [window] [lo hi]"""
import csv, sys
from collections import Counter
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr = rows[1]; data = rows[2:]
c = {n: i for i, n in enumerate(hdr)}
seen = set(); d2 = []
for r in data:
    k = r[c["Address"]]
    if k in seen: continue
    seen.add(k); d2.append(r)
def f(r, n):
    try: return float(r[c[n]])
    except Exception: return 0.0
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
tot = sum(f(r, "# Samples") for r in d2)
W = int(sys.argv[2]) if len(sys.argv) > 2 else 60
print("instructions", len(d2), "samples", tot)
agg = {s: sum(f(r, s) for r in d2) for s in stalls}
print("stall mix:", " ".join(f"{k[6:]}={v/tot*100:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
if len(sys.argv) > 4:
    lo, hi = int(sys.argv[3]), int(sys.argv[4])
    for i in range(lo, min(hi, len(d2))):
        r = d2[i]; s = f(r, "# Samples"); t = r[c["Source"]].strip()
        if s / tot * 100 >= 0.1 or any(m in t for m in ("LDG", "STG", "LDTM", "BAR", "SYNCS", "UTCHMMA", "UBLKCP", "ATOM", "RED")):
            top = sorted(stalls, key=lambda x: -f(r, x))[:2]
            print(i, f"{s/tot*100:5.2f}%", t[:84], " ".join(f"{x[6:]}={f(r,x):.0f}" for x in top if f(r, x) > 0))
else:
    for i in range(0, len(d2), W):
        blk = d2[i:i + W]
        s = sum(f(r, "# Samples") for r in blk)
        cn = Counter(m for r in blk for m in ("UTCHMMA", "BAR.SYNC", "LDG", "STG", "LDTM", "UBLKCP", "SYNCS", "MUFU", "STS", "LDS", "ATOM", "RED", "EXIT") if m in r[c["Source"]])
        if s / tot * 100 >= 0.3: print(f"{i:5d} {s/tot*100:5.1f}%  " + " ".join(f"{k}:{v}" for k, v in cn.items()))
