#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` export: total samples per stall reason and the hottest SASS lines.
   ncu -i X.ncu-rep --page source --csv > x.csv ; python tools/ncu_stalls.py x.csv [topN]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
print(rows[0][1][:200] if len(rows[0]) > 1 else "")
col = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = {h: 0 for h in stall_cols}
lines = []
for r in rows[hdr_i + 1:]:
    if len(r) < len(hdr): continue
    try: s = int(r[col["# Samples"]])
    except ValueError: continue
    for h in stall_cols:
        try: tot[h] += int(r[col[h]])
        except ValueError: pass
    lines.append((s, r[col["Source"]].strip(), {h: int(r[col[h]]) for h in stall_cols if r[col[h]].isdigit() and int(r[col[h]]) > 0}))
allS = sum(l[0] for l in lines)
print("total samples", allS)
for h, v in sorted(tot.items(), key=lambda kv: -kv[1])[:10]:
    print(f"  {h:28s} {v:8d} {100.0*v/max(allS,1):5.1f}%")
print("hottest lines:")
for i, (s, src, st) in enumerate(lines):
    pass
idx = sorted(range(len(lines)), key=lambda i: -lines[i][0])[:top]
for i in sorted(idx):
    s, src, st = lines[i]
    main = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    print(f"  [{i:4d}] {s:7d} {100.0*s/max(allS,1):5.1f}%  {src[:70]:70s} {main}")
