#!/usr/bin/env python
"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump per CUDA source line.
usage: ncu -i X.ncu-rep --page source --print-source cuda,sass --csv --launch-skip K --launch-count 1 > s.csv; python tools/ncu_lines.py s.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
fname = None; hdr = None; recs = []
for r in rows:
    if len(r) == 2 and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if len(r) > 5 and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < 10: continue
    if r[0] != "":  # a source line row (aggregated over its SASS)
        recs.append((fname, r))
n = len(hdr)  # source text may contain quotes/commas that split it: index the metric columns from the END of each row
ie = hdr.index("Instructions Executed") - n; te = hdr.index("Thread Instructions Executed") - n; sm = hdr.index("# Samples") - n
def num(x):
    try: return float(x)
    except ValueError: return 0.0
tot = sum(num(r[ie]) for _, r in recs)
tots = sum(num(r[sm]) for _, r in recs)
print(f"total warp-inst {tot:.0f}, samples {tots:.0f}")
for f, r in recs:
    try: v = float(r[ie]); s = float(r[sm])
    except ValueError: continue
    if v / tot * 100 >= thr or s / tots * 100 >= thr:
        print(f"{f[:22]:22s}:{r[0]:>4} inst {v/tot*100:5.1f}% thr/inst {float(r[te])/max(v,1):5.1f} samples {s/tots*100:5.1f}% | {r[1].strip()[:100]}")
