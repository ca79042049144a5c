#!/usr/bin/env python
"""Splits one kernel launch of an .ncu-rep (captured with --set full) at its BAR.SYNC instructions and
reports, per region, executed warp instructions and stall samples -- the per-phase cost of a kernel
whose phases are separated by block barriers (pf_fused_kernel: lookup | rounds | weigh).

    python profiles/ncu_phases.py rep.ncu-rep <kernel regex> [index]
"""
import collections
import csv
import io
import re
import subprocess
import sys

rep, kre = sys.argv[1], sys.argv[2]
skip = int(sys.argv[3]) if len(sys.argv) > 3 else 0
if rep.endswith(".csv"):        # a source page exported on the GPU box (run_r02_profiles.sh); the regex was applied there
    out = open(rep).read()
else:
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre],
                         capture_output=True, text=True).stdout
blocks = []
for r in csv.reader(io.StringIO(out)):
    if r and r[0] == "Kernel Name":
        blocks.append([])
    if blocks:
        blocks[-1].append(r)
rows = blocks[skip]
hdr = rows[1]
ix = {n: i for i, n in enumerate(hdr)}
regions = [dict(n=0, s=0, ops=collections.Counter(), stall=collections.Counter())]
warps = None
for r in rows[2:]:
    if len(r) < len(hdr) or r[ix["Source"]] == "Source":
        continue
    src = r[ix["Source"]].strip()
    m = re.match(r"(@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*(?:\.[A-Z0-9_]+)*)", src)
    if not m:
        continue
    op = m.group(2)
    n = int(r[ix["Instructions Executed"]] or 0)
    s = int(r[ix["# Samples"]] or 0)
    if warps is None:
        warps = n
    reg = regions[-1]
    reg["n"] += n
    reg["s"] += s
    reg["ops"][op.split(".")[0]] += n
    reg["stall"][src[:70]] += s
    if op.startswith("BAR"):
        regions.append(dict(n=0, s=0, ops=collections.Counter(), stall=collections.Counter()))
tot_n = sum(r["n"] for r in regions)
tot_s = sum(r["s"] for r in regions)
print("warps %d; %.1f instr/warp; %d samples" % (warps, tot_n / warps, tot_s))
for k, reg in enumerate(regions):
    if not reg["n"]:
        continue
    print("region %d: %7.1f instr/warp (%4.1f%%)  samples %6d (%4.1f%%)   top ops: %s" % (
        k, reg["n"] / warps, 100.0 * reg["n"] / tot_n, reg["s"], 100.0 * reg["s"] / max(tot_s, 1),
        ", ".join("%s %.0f" % (o, c / warps) for o, c in reg["ops"].most_common(6))))
    for src, s in reg["stall"].most_common(4):
        print("        %5d  %s" % (s, src))
