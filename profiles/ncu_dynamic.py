#!/usr/bin/env python
"""Dynamic instruction mix of one kernel launch from an .ncu-rep captured with --set full:
executed warp instructions per opcode and the stall-sample ranking (runs without a GPU).

    python profiles/ncu_dynamic.py rep.ncu-rep <kernel regex> [index among matching launches] [--list]
"""
import collections
import csv
import io
import re
import subprocess
import sys

rep, kre = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 and not sys.argv[3].startswith("--") else "0"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre],
                     capture_output=True, text=True).stdout
blocks = []
for r in csv.reader(io.StringIO(out)):
    if r and r[0] == "Kernel Name":
        blocks.append([])
    if blocks:
        blocks[-1].append(r)
print("%d matching launches; showing #%s" % (len(blocks), skip))
rows = blocks[int(skip)]
print(rows[0][1][:150])
hdr = rows[1]
ix = {n: i for i, n in enumerate(hdr)}
tot = collections.Counter()
samples = collections.Counter()
listing = []
for r in rows[2:]:
    if len(r) < len(hdr) or r[ix["Source"]] == "Source":
        continue
    src = r[ix["Source"]].strip()
    m = re.match(r"(@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*(?:\.[A-Z0-9_]+)*)", src)
    if not m:
        continue
    op = m.group(2).split(".")[0]
    n = int(r[ix["Instructions Executed"]] or 0)
    s = int(r[ix["# Samples"]] or 0)
    tot[op] += n
    samples[op] += s
    listing.append((n, s, src))
warps = listing[0][0]
total = sum(tot.values())
print("warps %d   executed warp-instructions %d   = %.1f per warp" % (warps, total, total / warps))
for op, n in tot.most_common(30):
    print("  %-8s %7.1f /warp   samples %6d" % (op, n / warps, samples[op]))
if "--list" in sys.argv:
    for n, s, src in listing:
        if n:
            print("%6.2f %5d  %s" % (n / warps, s, src[:100]))
