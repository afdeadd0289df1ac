#!/usr/bin/env python
"""Top CUDA source lines by executed warp instructions from an ncu report:
   ncu -i X.ncu-rep --page source --print-source cuda,sass --csv --kernel-name regex:K > f.csv; python tools/ncu_lines.py f.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
kern = None; hdr = None; agg = {}
for r in rows:
    if not r: continue
    if r[0] == "Function Name": kern = r[1][:60]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or kern is None: continue
    if r[0].isdigit() and len(r) > 8:
        try:
            ie = float(r[hdr.index("Instructions Executed")]); smp = float(r[hdr.index("# Samples")])
        except ValueError:
            continue
        d = agg.setdefault(kern, {})
        k = (int(r[0]), r[1].strip()[:110])
        a = d.setdefault(k, [0.0, 0.0]); a[0] += ie; a[1] += smp
for kern, d in agg.items():
    tot = sum(v[0] for v in d.values()); ts = sum(v[1] for v in d.values())
    print("== %s  warp-instr %.3g  samples %d" % (kern, tot, ts))
    for (ln, src), v in sorted(d.items(), key=lambda kv: -kv[1][0])[:top]:
        print("  %5.1f%% inst %5.1f%% smp  L%-4d %s" % (100 * v[0] / tot, 100 * v[1] / max(ts, 1), ln, src))
