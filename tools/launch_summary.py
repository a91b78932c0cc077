"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel:
    python tools/launch_summary.py launches.csv out_summary.csv
Per-launch times under ncu are cold-cache and serialised: the kernels' SHARES are what should agree with the bench."""
import csv
import re
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[hi]
kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
acc = defaultdict(list)
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    name = re.sub(r"\(.*", "", r[kn]).replace("b200q4::", "")
    v = float(r[mv].replace(",", ""))
    u = r[mu]
    v_us = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v
    acc[name].append(v_us)
tot = sum(sum(v) for v in acc.values())
with open(sys.argv[2], "w") as f:
    f.write("kernel,launches,mean_us,min_us,max_us,total_ms,share_of_gpu_time\n")
    for name, v in sorted(acc.items(), key=lambda kv: -sum(kv[1])):
        f.write(f"\"{name}\",{len(v)},{sum(v)/len(v):.2f},{min(v):.2f},{max(v):.2f},{sum(v)/1e3:.3f},{sum(v)/tot:.3f}\n")
print(open(sys.argv[2]).read())
