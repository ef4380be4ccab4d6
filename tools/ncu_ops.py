#!/usr/bin/env python
"""tools/ncu_ops.py report.ncu-rep kernel-regex [launch-skip] — per CUDA source line: share of the kernel's executed warp
instructions and the SASS opcodes behind it (finds conditionals that nvcc compiled into branches: BSSY / BRA / BSYNC)."""
import collections, csv, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", "regex:" + rx,
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; cur = None; fname = ""
per = collections.defaultdict(collections.Counter); src = {}; lanes = collections.defaultdict(lambda: [0.0, 0.0]); tot = 0.0
for r in rows:
    if len(r) == 2 and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if len(r) == 2 and r[0] == "Function Name": print(r[1][:110]); continue
    if len(r) > 5 and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < 10: continue
    n = len(hdr)
    if r[0] != "":
        cur = (fname, r[0]); src[cur] = r[1].strip()[:70]
    else:
        try:
            v = float(r[hdr.index("Instructions Executed") - n]); t = float(r[hdr.index("Thread Instructions Executed") - n])
        except ValueError:
            continue
        w = r[3].strip().split()
        if not w: continue
        op = (w[1] if w[0].startswith("@") else w[0]).split(".")[0]
        per[cur][op] += v; lanes[cur][0] += t; lanes[cur][1] += v; tot += v
thr = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
for s, k, c in sorted(((sum(c.values()), k, c) for k, c in per.items()), reverse=True):
    if s / tot * 100 < thr: break
    top = ", ".join(f"{o}:{v / tot * 100:.1f}" for o, v in c.most_common(5))
    print(f"{s / tot * 100:5.1f}% {lanes[k][0] / max(lanes[k][1], 1):5.1f} lanes {k[0][:20]:20s}:{k[1]:>4} | {top} | {src[k]}")
