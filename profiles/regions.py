#!/usr/bin/env python
"""Attribute executed warp instructions of the AF scan kernel to regions of vcfx_kernels.cuh.
usage: regions.py src.csv <windows> <lines> [kernel-substring]   (src.csv from
       ncu -i X.ncu-rep --page source --print-source cuda,sass --csv)"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
nwin = float(sys.argv[2]); nlines = float(sys.argv[3]); want = sys.argv[4] if len(sys.argv) > 4 else "(int)1"
fn = None; hdr = None; fpath = None; agg = {}; other = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": fpath = r[1]; continue
    if r[0] == "Function Name": fn = r[1]; hdr = None; continue
    if r[0] == "Line No": hdr = r; iI = r.index("Instructions Executed"); iT = r.index("Thread Instructions Executed"); continue
    if fn is None or hdr is None or want not in fn: continue
    if r[0].isdigit() and r[0] != "0":
        try: v = (int(r[iI] or 0), int(r[iT] or 0))
        except ValueError: continue
        d = agg if fpath.endswith("vcfx_kernels.cuh") else other
        k = int(r[0]) if d is agg else fpath.split("/")[-1]
        a = d.get(k, (0, 0)); d[k] = (a[0] + v[0], a[1] + v[1])
tot = sum(v[0] for v in agg.values()) + sum(v[0] for v in other.values())
print(f"total {tot/nwin:.1f} warp-instr per 512-B window, {tot/nlines:.0f} per line")
lines = open(__file__.rsplit("/", 2)[0] + "/vcfx_b200/csrc/vcfx_kernels.cuh").read().split("\n")
def find(txt):
    for i, l in enumerate(lines, 1):
        if txt in l: return i
marks = [("helpers", 1), ("slow parsers", find("exact scalar parsers")), ("sample_reg/generic", find("per-lane sample parsing from registers")),
         ("lattice fn", find("Lattice check of one lane")), ("kernel prologue", find("K1: the fused scan")),
         ("first line scan", find("first line start in [a, b)")), ("line start", find("every line that starts in the tile")),
         ("header phase", find("header phase: rank tabs")), ("header decisions", find("decisions that need only the header")),
         ("sample phase setup", find("================= sample phase")), ("T1 check", find("if (t1_on && prev_ok)")),
         ("exact window path", find("uint32_t packed = 0;")), ("find-eol loop", find("no (more) per-sample work")),
         ("end of line", find("end of line: [ls, ee)")), ("K2", find("K2a: exclusive scans"))]
marks = [(n, l) for n, l in marks if l]
for i, (n, l) in enumerate(marks):
    hi = marks[i + 1][1] - 1 if i + 1 < len(marks) else 99999
    ins = sum(v[0] for k, v in agg.items() if l <= k <= hi); th = sum(v[1] for k, v in agg.items() if l <= k <= hi)
    print(f"{n:22s} L{l:4d}-{hi:5d} {100*ins/tot:5.1f}% {ins/nwin:6.1f}/win {ins/nlines:7.0f}/line  act {th/max(ins,1):4.1f}")
for k, v in sorted(other.items(), key=lambda kv: -kv[1][0]):
    print(f"{'[' + k + ']':34s} {100*v[0]/tot:5.1f}% {v[0]/nwin:6.1f}/win {v[0]/nlines:7.0f}/line  act {v[1]/max(v[0],1):4.1f}")
