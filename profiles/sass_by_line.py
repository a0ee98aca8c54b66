#!/usr/bin/env python
"""Executed warp instructions of one kernel per CUDA source line, from SASS-level counts.
usage: sass_by_line.py <ncu sass csv> <cubin> <mangled kernel substring> <ncu kernel substring> <lines in the input> [min per line]
  ncu -i X.ncu-rep --page source --print-source sass --csv > sass.csv ; cuobjdump -xelf all lib.so
The ncu SASS page lists the kernel's instructions in address order; nvdisasm -g gives the source
line of each; the two are joined by position (the cuda,sass page of ncu elides instructions)."""
import csv, subprocess, sys, re, collections
sass_csv, cubin, mangled, want, nlines = sys.argv[1:6]
nlines = float(nlines); thr = float(sys.argv[6]) if len(sys.argv) > 6 else 2.0
rows = list(csv.reader(open(sass_csv)))
kern = []; cur = None
for r in rows:
    if not r: continue
    if r[0] == "Kernel Name": cur = {"name": r[1], "ins": []}; kern.append(cur); continue
    if r[0] == "Address": iI = r.index("Instructions Executed"); iS = r.index("Source"); iW = r.index("Warp Stall Sampling (All Samples)"); continue
    if cur is not None:
        try: cur["ins"].append((r[iS].strip(), int(r[iI] or 0), int(r[iW] or 0)))
        except (ValueError, IndexError): pass
k = [k for k in kern if want in k["name"]][0]["ins"]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.split("\n")
ins = []; on = False; f = l = None
for ln in dis:
    if ln.startswith(".text.") and ln.endswith(":"): on = mangled in ln and "$" not in ln.split(mangled, 1)[1]; continue
    if ln.startswith("//-----") : on = False if on and ins else on
    if not on: continue
    m = re.match(r'\s*//## File "(.*)", line (\d+)', ln)
    if m: f, l = m.group(1).split("/")[-1], int(m.group(2)); continue
    m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", ln)
    if m: ins.append((f, l, m.group(2)))
n = min(len(ins), len(k))
print(f"# {len(ins)} instructions in the cubin, {len(k)} in the profile; total {sum(x[1] for x in k) / nlines:.0f} executed per line, {sum(x[2] for x in k)} stall samples")
agg = collections.OrderedDict()
for i in range(n):
    key = (ins[i][0], ins[i][1]); a = agg.setdefault(key, [0, 0]); a[0] += k[i][1]; a[1] += k[i][2]
src = {}
try:
    for i, t in enumerate(open(__file__.rsplit("/", 2)[0] + "/vcfx_b200/csrc/vcfx_kernels.cuh"), 1): src[i] = t.strip()
except OSError: pass
tot_s = sum(x[2] for x in k) or 1
for (f, l), (c, s) in sorted(agg.items(), key=lambda kv: (kv[0][0] != "vcfx_kernels.cuh", kv[0][1])):
    if c / nlines >= thr: print(f"{f[:20]:20s} L{l:<5d} {c / nlines:7.1f}/line {100 * s / tot_s:5.1f}% stall  {src.get(l, '')[:100] if f == 'vcfx_kernels.cuh' else ''}")
