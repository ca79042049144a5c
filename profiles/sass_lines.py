#!/usr/bin/env python
"""Static SASS instruction count per SOURCE LINE of one kernel (needs -lineinfo; runs without a GPU).

    python profiles/sass_lines.py <object-name e.g. pf_step_mvn> <mangled-kernel-substring> [top]
"""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "cusmc_b200", "libcusmc_b200.so")
obj, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
with tempfile.TemporaryDirectory() as tmp:
    subprocess.run(["cuobjdump", "-xelf", obj + ".sm_100a.cubin", so], cwd=tmp, capture_output=True)
    sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, obj + ".sm_100a.cubin")],
                          capture_output=True, text=True).stdout
inside, cur = False, ("?", 0)
per_line = collections.Counter()
ops = collections.defaultdict(collections.Counter)
for line in sass.splitlines():
    if line.startswith("//--------------------- .text."):
        inside = pat in line
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m:
        per_line[cur] += 1
        ops[cur][m.group(2)] += 1
total = sum(per_line.values())
print("total static instructions:", total)
for (f, l), n in per_line.most_common(top):
    print("%5d  %-22s:%-5d %s" % (n, f, l, " ".join("%s*%d" % kv for kv in ops[(f, l)].most_common(6))))
