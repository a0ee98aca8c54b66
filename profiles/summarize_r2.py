#!/usr/bin/env python
"""Round-2 summaries from ncu captures.
usage: summarize_r2.py <out.json> <report.ncu-rep> [label ...]     selected metrics of every profiled launch, in launch order
       summarize_r2.py --launches <out.txt> <launches.csv>         per-kernel share of a `--metrics gpu__time_duration.sum` launch list"""
import collections, csv, json, subprocess, sys

WANT = ["launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def full(out, rep, labels):
    raw = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
    h, units = raw[0], raw[1]
    res = []
    for n, r in enumerate(raw[2:]):
        d = {"launch": n, "label": labels[n] if n < len(labels) else "", "Kernel Name": r[h.index("Kernel Name")]}
        for w in WANT:
            if w in h:
                d[w] = (r[h.index(w)] + " " + units[h.index(w)]).strip()
        res.append(d)
    json.dump(res, open(out, "w"), indent=1)
    for d in res:
        print(d["launch"], d["label"], d["Kernel Name"][:40], d.get("gpu__time_duration.sum"), d.get("dram__bytes_read.sum"), d.get("dram__bytes_write.sum"), d.get("smsp__inst_executed.sum"))


def launches(out, path):
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("==")) if r]
    hdr = rows[0]; iK = hdr.index("Kernel Name"); iV = hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[iV].replace(",", ""))
        except ValueError:
            continue
        a = agg.setdefault(r[iK].split("(")[0], [0, 0.0]); a[0] += 1; a[1] += v / 1e3
    tot = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write("# ncu launch list (gpu__time_duration.sum); per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes\n")
        f.write(f"{'kernel':70s} {'launches':>8s} {'total_us':>10s} {'avg_us':>9s} {'share':>7s}\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k[:70]:70s} {n:8d} {t:10.1f} {t / n:9.1f} {100 * t / tot:6.1f}%\n")
    print(open(out).read())


if __name__ == "__main__":
    if sys.argv[1] == "--launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[1], sys.argv[2], sys.argv[3:])
