"""Per-kernel share of the step from an `ncu --metrics gpu__time_duration.sum --csv` launch list.  usage: launch_summary.py list.csv"""
import csv, sys
from collections import defaultdict
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
tot = defaultdict(float); cnt = defaultdict(int)
for r in rows[1:]:
    if len(r) <= vi: continue
    v = float(r[vi].replace(",", "")) * {"ns": 1.0, "us": 1e3, "ms": 1e6}.get(r[ui], 1.0)
    tot[r[ki]] += v; cnt[r[ki]] += 1
s = sum(tot.values())
print(f"{sum(cnt.values())} launches of this library's kernels, {s/1e6:.3f} ms in total (unit ns; cold-cache serialised times: compare shares)")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{v/s*100:6.2f}%  n={cnt[k]:4d}  total={v:14.0f}  {k[:120]}")
