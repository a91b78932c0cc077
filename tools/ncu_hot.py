"""Development aid: top stalled SASS instructions of an ncu --page source --csv dump.
    ncu -i X.ncu-rep --page source --csv > src.csv ; python tools/ncu_hot.py src.csv [N]
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
for n, r in enumerate(rows[hi + 1:]):
    if len(r) < len(hdr):
        continue
    data.append((int(r[col["# Samples"]] or 0), n, r))
tot = sum(d[0] for d in data)
print("total samples", tot)
agg = {s: sum(int(d[2][col[s]] or 0) for d in data) for s in stalls}
print("by reason:", sorted(((v, k) for k, v in agg.items() if v), reverse=True)[:8])
for smp, n, r in sorted(data, reverse=True)[:topn]:
    top = sorted(((int(r[col[s]] or 0), s[6:]) for s in stalls), reverse=True)[:3]
    print(f"{smp:7d} {100.0 * smp / max(tot, 1):5.1f}%  #{n:5d} exec={r[col['Instructions Executed']]:>9}  {r[col['Source']].strip()[:70]:70s} {top}")
