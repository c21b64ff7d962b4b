"""Top SASS lines of an `ncu --page source --csv` export by stall samples, plus shared-memory excess wavefronts."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; data = rows[2:]
c = {n: i for i, n in enumerate(hdr)}
def f(r, n):
    try: return float(r[c[n]])
    except Exception: return 0.0
tot = sum(f(r, "# Samples") for r in data)
print("total samples", tot, "instructions", len(data))
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
agg = {s: sum(f(r, s) for r in data) for s in stalls}
print("stall mix:", " ".join(f"{k[6:]}={v/tot*100:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
print("--- top lines by samples")
for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    top = sorted(stalls, key=lambda s: -f(r, s))[:2]
    print(f"{f(r,'# Samples')/tot*100:5.1f}%  {r[c['Source']].strip()[:90]:90s} {' '.join(f'{s[6:]}={f(r,s):.0f}' for s in top)}")
print("--- shared excessive wavefronts")
for r in sorted(data, key=lambda r: -f(r, "L1 Wavefronts Shared Excessive"))[:8]:
    if f(r, "L1 Wavefronts Shared Excessive") > 0:
        print(f"{f(r,'L1 Wavefronts Shared Excessive'):12.0f} of {f(r,'L1 Wavefronts Shared'):12.0f}  {r[c['Source']].strip()[:90]}")
print("--- global excessive sectors")
for r in sorted(data, key=lambda r: -f(r, "L2 Theoretical Sectors Global Excessive"))[:6]:
    if f(r, "L2 Theoretical Sectors Global Excessive") > 0:
        print(f"{f(r,'L2 Theoretical Sectors Global Excessive'):12.0f} of {f(r,'L2 Theoretical Sectors Global'):12.0f}  {r[c['Source']].strip()[:90]}")
