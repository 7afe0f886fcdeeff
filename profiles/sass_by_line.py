#!/usr/bin/env python3
"""Per-source-line instruction counts and stall samples of one kernel: joins `ncu --page source --csv` (SASS rows with
counters) with `nvdisasm -g` line info of the same cubin by instruction offset.
usage: sass_by_line.py <ncu_source.csv> <nvdisasm_function.sass> [top] [kernel block]"""
import csv
import re
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
which = int(sys.argv[4]) if len(sys.argv) > 4 else 0  # the file holds one block per profiled kernel
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
rows = rows[starts[which]:starts[which + 1]]
print(rows[0][1][:100])
hdr = rows[1]
iA, iS, iSm, iEx, iTh = (hdr.index(k) for k in ("Address", "Source", "# Samples", "Instructions Executed", "Thread Instructions Executed"))
inst = [(r[iS].strip(), int(r[iSm] or 0), int(r[iEx] or 0), int(r[iTh] or 0)) for r in rows[2:] if len(r) > iTh]
line = None
lines = []
for l in open(sys.argv[2]):
    m = re.search(r'//## File "(.*?)", line (\d+)', l)
    if m:
        line = "%s:%s" % (m.group(1).split("/")[-1], m.group(2))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        lines.append((int(m.group(1), 16), line, m.group(2).strip()))
assert len(lines) >= len(inst), (len(lines), len(inst))
agg = defaultdict(lambda: [0, 0, 0])
tot = [0, 0, 0]
for (src, sm, ex, th), (off, ln, txt) in zip(inst, lines):
    a = agg[ln]
    a[0] += sm; a[1] += ex; a[2] += th
    tot[0] += sm; tot[1] += ex; tot[2] += th
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
print("total samples %d, warp instructions %d, thread instructions %d" % tuple(tot))
print("%28s %8s %6s %12s %6s %5s" % ("line", "samples", "%", "warp-inst", "%", "thr/w"))
for ln, (sm, ex, th) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%28s %8d %6.2f %12d %6.2f %5.1f" % (ln, sm, 100.0 * sm / max(tot[0], 1), ex, 100.0 * ex / max(tot[1], 1), th / max(ex, 1)))
