#!/usr/bin/env python
"""Static SASS statistics of the library's kernels (runs without a GPU):

    python profiles/sass_stats.py [regex]            # instruction count per kernel
    python profiles/sass_stats.py regex --ops        # + opcode histogram

Straight-line kernels execute (almost) every instruction once per thread, so the static count is a
good stand-in for ncu's `smsp__inst_executed` per warp while iterating on CPU."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "cusmc_b200", "libcusmc_b200.so")
pat = re.compile(sys.argv[1]) if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else re.compile(".")
show_ops = "--ops" in sys.argv
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
demangle = {}
cur, stats = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        stats[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur:
        stats[cur][m.group(2)] += 1
names = list(stats)
dm = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
for mangled, nice in zip(names, dm):
    nice = re.sub(r"\(anonymous namespace\)::", "", nice)
    nice = nice.split("(")[0]
    if not pat.search(nice):
        continue
    c = stats[mangled]
    print("%-60s %6d instr  DFMA %4d  FFMA/FMUL/FADD %4d  IMAD %4d" %
          (nice[:60], sum(c.values()), c["DFMA"] + c["DMUL"] + c["DADD"], c["FFMA"] + c["FMUL"] + c["FADD"], c["IMAD"]))
    if show_ops:
        print("    " + "  ".join("%s %d" % kv for kv in c.most_common(24)))
