"""SASS opcode census of libb200q4.so per kernel: the mnemonics that prove which hardware paths the shipped code uses
(tcgen05: UTCQMMA / UTCHMMA / UTCBAR, TMEM loads: LDTM, bulk-copy engine: UBLKCP / UBLKPF, legacy tensor pipe: IMMA / HMMA,
mbarriers: SYNCS, named barriers: BAR.ARV / BAR.SYNC).   python tools/sass_census.py > profiles/rNN_sass_census.txt"""
import collections
import re
import subprocess
import sys
from pathlib import Path

lib = Path(__file__).resolve().parents[1] / "vsim_b200" / "libb200q4.so"
sass = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True).stdout
WATCH = ["UTCQMMA", "UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UBLKPF", "UTMALDG", "IMMA", "HMMA", "SYNCS", "BAR.ARV", "BAR.SYNC",
         "REDG", "ATOMG", "LDL", "STL", "DADD", "DMUL", "MUFU"]
per = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur).replace("b200q4::", "").replace("(anonymous namespace)::", "")
        per.setdefault(cur, collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        per[cur]["#instr"] += 1
        for w in WATCH:
            if op.startswith(w):
                per[cur][w] += 1
print(f"SASS census of {lib.name} (cuobjdump -sass, sm_100a).  Columns: instructions, then the watched mnemonics that occur.")
tot = collections.Counter()
for k, c in per.items():
    tot.update(c)
    print(f"{k:70s} {c['#instr']:6d}  " + " ".join(f"{w}={c[w]}" for w in WATCH if c[w]))
print(f"{'TOTAL':70s} {tot['#instr']:6d}  " + " ".join(f"{w}={tot[w]}" for w in WATCH if tot[w]))
