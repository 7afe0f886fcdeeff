#!/usr/bin/env python3
"""Turns gpurun_out/ ncu outputs into the committed summaries under profiles/.

usage: python profiles/summarize.py <launches.csv> <prof.ncu-rep> <out.md> [title]
"""
import collections
import csv
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum"]


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1000 if r[ui] == "ns" else (v * 1000 if r[ui] == "ms" else v)
        a = agg.setdefault(r[ki].split("(")[0], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    out = ["| kernel | launches | total us | us / launch | share |", "|---|---:|---:|---:|---:|"]
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("| `%s` | %d | %.1f | %.2f | %.1f %% |" % (k, c, t, t / c, 100 * t / tot))
    return out


def raw(path):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [(m, hdr.index(m)) for m in METRICS if m in hdr]
    ki = hdr.index("Kernel Name")
    out = ["| kernel | " + " | ".join("%s [%s]" % (m, units[i]) for m, i in idx) + " |", "|---|" + "---:|" * len(idx)]
    for r in rows[2:]:
        out.append("| `%s` | " % r[ki].split("(")[0] + " | ".join(r[i] for _, i in idx) + " |")
    return out


if __name__ == "__main__":
    lpath, rpath, opath = sys.argv[1:4]
    title = sys.argv[4] if len(sys.argv) > 4 else "ncu summary"
    lines = ["# " + title, ""]
    if lpath != "-":
        lines += ["## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`; cold-cache, serialised: compare shares)", ""] + launches(lpath) + [""]
    if rpath != "-":
        lines += ["## Full captures (`ncu --set full --clock-control none --import-source on`), raw page extract", ""] + raw(rpath) + [""]
    open(opath, "w").write("\n".join(lines))
    print("\n".join(lines))
