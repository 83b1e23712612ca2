#!/usr/bin/env python
"""Summarise an ncu report (source page): per-source-line stall samples, hottest SASS, opcode mix.
Usage: ncu_hot.py report.ncu-rep [kernel-index] [top]"""
import csv
import subprocess
import sys
from collections import Counter, defaultdict

rep = sys.argv[1]
kidx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
starts.append(len(rows))
hdr = rows[starts[kidx] + 1]
body = rows[starts[kidx] + 2:starts[kidx + 1]]
print(rows[starts[kidx]][1][:150])
iS, iSrc, iEx = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[iS]) for r in body)
print("total samples", tot, "sass rows", len(body))
agg = Counter()
for r in body:
    for i in stall_cols:
        agg[hdr[i][6:]] += int(r[i])
print("stall totals:", ", ".join(f"{k} {100 * v / max(1, sum(agg.values())):.1f}%" for k, v in agg.most_common(10)))
for r in sorted(body, key=lambda r: -int(r[iS]))[:top]:
    st = sorted(((hdr[i][6:], int(r[i])) for i in stall_cols if int(r[i]) > 0), key=lambda kv: -kv[1])[:3]
    print(f"{int(r[iS]):6d} {100 * int(r[iS]) / tot:5.1f}% ex={r[iEx]:>9} {r[iSrc].strip()[:80]:80s} {st}")
c, s = Counter(), Counter()
for r in body:
    toks = r[iSrc].split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    op = op.split(".")[0]
    c[op] += int(r[iEx])
    s[op] += int(r[iS])
te = sum(c.values())
print("opcode mix (executed warp-instr share / sample share):")
for op, n in c.most_common(18):
    print(f"  {op:10s} {100 * n / te:5.1f}%  {100 * s[op] / tot:5.1f}%")
