#!/usr/bin/env python
"""Print the kernels of one bench step from an `ncu --metrics gpu__time_duration.sum --csv` launch list (one step = the
launches between two consecutive stem kernels).  Usage: step_launches.py launches.csv [step-index]"""
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[h]
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
L = [(r[ik], float(r[iv].replace(",", "")) / 1e3) for r in rows[h + 1:] if len(r) > iv and r[iv] not in ("", "Metric Value")]
idx = [i for i, (k, _) in enumerate(L) if "stem" in k]
step = int(sys.argv[2]) if len(sys.argv) > 2 else 3
a, b = idx[step], idx[step + 1]
tot = 0.0
for k, t in L[a:b]:
    tot += t
    print(f"{t:8.1f} us  {re.sub(r'[(].*', '', k).replace('void ', '').replace('<unnamed>::', '')[:100]}")
print(f"{tot:8.1f} us in {b - a} kernels")
