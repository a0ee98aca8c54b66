#!/usr/bin/env python
"""SASS of the loops the numbers depend on, from the built library (no GPU needed).
usage: sass_loops.py <libvcfx_cuda.so> > r2_sass_loops.txt
For each (kernel, source-line range) below: the SASS instructions nvdisasm attributes to those lines of
vcfx_kernels.cuh, in address order, with an opcode histogram."""
import collections, os, re, subprocess, sys, tempfile

lib = os.path.abspath(sys.argv[1])
src = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "vcfx_b200", "csrc", "vcfx_kernels.cuh")
text = open(src).read().split("\n")


def find(pattern, start=0):
    for i in range(start, len(text)):
        if pattern in text[i]:
            return i + 1
    raise SystemExit(f"pattern not found: {pattern}")


t1a = find("for (; it < ITMAX; ++it) {")
t1b = find("#undef VCFX_T1_ROUND")
dga = find("__device__ __forceinline__ void digits_window(")
dgb = find("template <int OP>", dga)
mka = find("__device__ __forceinline__ bool multikey_window(")
mkb = find("return odd;", mka)
bka = find("__device__ __forceinline__ void warp_flush_smem_bulk(")
bkb = find("#endif", bka)
sections = [("vcfx_scan_kernelILi1ELi0E", t1a, t1b, "tier-1 steady rounds (allele_freq_calc, lattice kernel)"),
            ("vcfx_scan_kernelILi1ELi1E", dga, dgb, "digit path: one window (allele_freq_calc, general kernel)"),
            ("vcfx_scan_kernelILi1ELi1E", mka, mkb, "skip-ahead loop: one window (allele_freq_calc, general kernel)"),
            ("vcfx_scan_kernelILi4ELi0E", bka, bkb, "allele_counter: bulk-copy flush of staged rows (UBLKCP = cp.async.bulk)")]
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=td, capture_output=True)
    cubin = [os.path.join(td, f) for f in os.listdir(td) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.split("\n")
for mangled, lo, hi, title in sections:
    on = False; line = None; rows = []
    for ln in dis:
        if ln.startswith(".text.") and ln.endswith(":"):
            on = mangled in ln
            continue
        if not on:
            continue
        m = re.match(r'\s*//## File "(.*)", line (\d+)', ln)
        if m:
            line = int(m.group(2)) if m.group(1).endswith("vcfx_kernels.cuh") else None
            continue
        m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", ln)
        if m and line and lo <= line <= hi:
            rows.append((m.group(1), line, m.group(2)))
    hist = collections.Counter((r[2].split()[1] if r[2].startswith("@") else r[2].split()[0]).split(".")[0] for r in rows)
    print(f"## {title}\n## kernel {mangled}, vcfx_kernels.cuh lines {lo}-{hi}: {len(rows)} SASS instructions")
    print("## " + ", ".join(f"{k} {v}" for k, v in hist.most_common()))
    for a, l, op in rows:
        print(f"  /*{a}*/ L{l:<5d} {op}")
    print()
