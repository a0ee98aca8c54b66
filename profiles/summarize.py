#!/usr/bin/env python
"""Turn the ncu captures of a round into the tracked summaries under profiles/.
usage: summarize.py <round tag> <launches.csv> <full .ncu-rep> <libvcfx_cuda.so> <lines in input>
  launches.csv : ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file ... python bench.py ...
  .ncu-rep     : ncu --set full --clock-control none --import-source on -k regex:vcfx_scan_kernel ... (AF and VC instantiations)
writes <tag>_launches_c2.csv, <tag>_launches_c2_summary.txt, <tag>_ncu_full_c2.json, <tag>_lines_af.txt"""
import collections, csv, json, os, shutil, subprocess, sys, tempfile
tag, launches, rep, lib, nlines = sys.argv[1:6]
here = os.path.dirname(os.path.abspath(__file__))

# 1. launch list -> shares
rows = [r for r in csv.reader(l for l in open(launches) if not l.startswith("==")) if r]
hdr = rows[0]; iK = hdr.index("Kernel Name"); iV = hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    try: v = float(r[iV].replace(",", ""))
    except ValueError: continue
    k = r[iK].split("(")[0]
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v / 1e3     # ns -> us
tot = sum(a[1] for a in agg.values())
shutil.copy(launches, f"{here}/{tag}_launches_c2.csv")
with open(f"{here}/{tag}_launches_c2_summary.txt", "w") as f:
    f.write("# ncu launch list, bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline (C2, 4.30 GB resident), first 60 launches\n")
    f.write("# per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes\n")
    f.write(f"{'kernel':62s} {'launches':>8s} {'total_us':>10s} {'avg_us':>9s} {'share':>7s}\n")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k[:62]:62s} {n:8d} {t:10.1f} {t / n:9.1f} {100 * t / tot:6.1f}%\n")

# 2. selected metrics of the full capture
want = ["launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
raw = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
h, units = raw[0], raw[1]
out = []
for r in raw[2:]:
    d = {"Kernel Name": r[h.index("Kernel Name")]}
    for w in want:
        if w in h: d[w] = (r[h.index(w)] + " " + units[h.index(w)]).strip()
    out.append(d)
json.dump(out, open(f"{here}/{tag}_ncu_full_c2.json", "w"), indent=1)

# 3. executed instructions per source line of the AF kernel
with tempfile.TemporaryDirectory() as td:
    sass = os.path.join(td, "sass.csv")
    open(sass, "w").write(subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv"], capture_output=True, text=True).stdout)
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=td, capture_output=True)
    cubin = [os.path.join(td, f) for f in os.listdir(td) if f.endswith(".cubin")][0]
    txt = subprocess.run([sys.executable, f"{here}/sass_by_line.py", sass, cubin, "vcfx_scan_kernelILi1E", "(int)1", nlines, "4"], capture_output=True, text=True).stdout
    open(f"{here}/{tag}_lines_af.txt", "w").write(txt)
print(open(f"{here}/{tag}_launches_c2_summary.txt").read())
for d in out: print(d["Kernel Name"], d.get("gpu__time_duration.sum"), d.get("dram__bytes_read.sum"), d.get("dram__bytes_write.sum"), d.get("smsp__inst_executed.sum"))
