#!/usr/bin/env python
"""Rank CUDA source lines of one kernel by executed warp instructions.
usage: ncu -i X.ncu-rep --page source --print-source cuda,sass --csv > src.csv ; src_hot.py src.csv '<kernel substring>' [N]
"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
want = sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
fn = None; hdr = None; agg = {}; src = {}
for r in rows:
    if not r: continue
    if r[0] == "Function Name": fn = r[1]; hdr = None; continue
    if r[0] == "Line No": hdr = r; iI = r.index("Instructions Executed"); iW = r.index("Warp Stall Sampling (All Samples)"); continue
    if fn is None or want not in fn or hdr is None: continue
    if r[0].isdigit() and r[0] != "0":
        cur = int(r[0]); src[cur] = r[1]
        try: agg[cur] = (int(r[iI] or 0), int(r[iW] or 0))
        except ValueError: pass
tot = sum(v[0] for v in agg.values()) or 1; tw = sum(v[1] for v in agg.values()) or 1
print(f"{want}: {tot} warp instructions over {len(agg)} source lines; {tw} stall samples")
for ln, (i, w) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*i/tot:5.1f}% inst {100*w/tw:5.1f}% stall  L{ln:<4} {src[ln].strip()[:120]}")
