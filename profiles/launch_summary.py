#!/usr/bin/env python
"""Per-kernel launch counts and durations from an `ncu --metrics gpu__time_duration.sum --csv` log.
usage: python profiles/launch_summary.py launches.csv "title" > summary.md"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
title = sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]
hdr, agg = None, collections.OrderedDict()
for r in rows:
    if r and r[0] == "ID":
        hdr = {n: i for i, n in enumerate(r)}
        continue
    if not hdr or len(r) < len(hdr):
        continue
    try:
        v = float(r[hdr["Metric Value"]].replace(",", ""))
    except ValueError:
        continue
    unit = r[hdr["Metric Unit"]]
    v = v / 1e3 if unit in ("nsecond", "ns") else v * 1e3 if unit in ("msecond", "ms") else v
    name = re.sub(r"\(.*", "", r[hdr["Kernel Name"]])[:70]
    agg.setdefault((name, r[hdr["Grid Size"]]), []).append(v)
lib = ("at::", "cutlass", "potrf", "gemv", "internal::kernel")
print("# %s\n" % title)
print("ncu serialises launches and flushes caches between them: the SHARE of a step is what these numbers give.\n")
print("| kernel | grid | launches | mean us | min us |\n|---|---|---|---|---|")
other = 0
for (name, grid), v in agg.items():
    if any(k in name for k in lib):
        other += 1
        continue
    print("| %s | %s | %d | %.2f | %.2f |" % (name, grid, len(v), sum(v) / len(v), min(v)))
print("\n%d further kernel names in the raw list are torch / cuSOLVER / cuBLAS launches that synthesise the inputs "
      "(randn, batched Cholesky of the per-chain covariances) outside every timed region." % other)
