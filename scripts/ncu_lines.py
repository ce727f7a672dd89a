#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export by source line.

usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > x.csv; ncu_lines.py x.csv [N]
"""
import csv, sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None
agg = defaultdict(lambda: {"src": "", "smp": 0, "ins": 0, "n": 0, "st": defaultdict(int)})
fname = ""
for r in rows:
    if r and r[0] == "File Path":
        fname = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No":
        hdr = r
        ix = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples")
        stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr is None or len(r) <= ix:
        continue
    try:
        ln = int(r[0]); n = int(r[ix]); s = int(r[isamp])
    except ValueError:
        continue
    a = agg[(fname, ln)]
    a["src"] = r[1]; a["smp"] += s; a["ins"] += n; a["n"] += 1
    for i in stall:
        a["st"][hdr[i][6:]] += int(r[i] or 0)
ts = sum(a["smp"] for a in agg.values()); ti = sum(a["ins"] for a in agg.values())
print(f"samples {ts}  warp-instructions {ti/1e9:.3f}e9")
tot_st = defaultdict(int)
for a in agg.values():
    for k, v in a["st"].items(): tot_st[k] += v
print("stalls:", {k: round(v / ts, 3) for k, v in sorted(tot_st.items(), key=lambda kv: -kv[1])[:9]})
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1]["smp"])[:top]:
    st = sorted(a["st"].items(), key=lambda kv: -kv[1])[:3]
    print(f"{f[:14]:14s}{ln:5d} {a['smp']/ts*100:5.1f}%smp {a['ins']/ti*100:5.1f}%ins {a['n']:4d}sass "
          f"{a['src'].strip()[:80]:80s} {[(k, v) for k, v in st]}")
