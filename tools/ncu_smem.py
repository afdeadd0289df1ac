#!/usr/bin/env python
"""Shared-memory wavefronts per CUDA source line from an ncu report's source page (csv):
   ncu -i X.ncu-rep --page source --print-source cuda,sass --csv > f.csv; python tools/ncu_smem.py f.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
kern = None; hdr = None; agg = {}
for r in rows:
    if not r: continue
    if r[0] == "Function Name": kern = r[1][:60]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or kern is None: continue
    if r[0].isdigit() and len(r) > 20:
        try:
            wf = float(r[hdr.index("L1 Wavefronts Shared")] or 0); ideal = float(r[hdr.index("L1 Wavefronts Shared Ideal")] or 0)
            ie = float(r[hdr.index("Instructions Executed")] or 0)
        except ValueError:
            continue
        d = agg.setdefault(kern, {})
        k = (int(r[0]), r[1].strip()[:100])
        a = d.setdefault(k, [0.0, 0.0, 0.0]); a[0] += wf; a[1] += ideal; a[2] += ie
for kern, d in agg.items():
    tot = sum(v[0] for v in d.values()); ti = sum(v[1] for v in d.values())
    print("== %s  shared wavefronts %.3g (ideal %.3g)" % (kern, tot, ti))
    for (ln, src), v in sorted(d.items(), key=lambda kv: -kv[1][0])[:top]:
        print("  %5.1f%% wf  x%4.2f of ideal  L%-4d %s" % (100 * v[0] / max(tot, 1), v[0] / max(v[1], 1), ln, src))
