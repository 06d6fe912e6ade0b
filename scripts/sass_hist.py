#!/usr/bin/env python
"""Static opcode histogram of one kernel in a .so / cubin (cuobjdump -sass): a CPU-side proxy for
instructions per sample before spending GPU time.  usage: sass_hist.py <file> <kernel-substring> [samples_per_body]"""
import collections
import re
import subprocess
import sys

f, pat = sys.argv[1], sys.argv[2]
per = float(sys.argv[3]) if len(sys.argv) > 3 else None
out = subprocess.run(["cuobjdump", "-sass", f], capture_output=True, text=True).stdout
cur, hist, total = None, collections.Counter(), 0
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        continue
    if cur is None or pat not in cur:
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m:
        op = m.group(1)
        if op == "NOP":
            continue
        hist[op.split(".")[0] + ("." + op.split(".")[1] if op.startswith(("LDS", "STS", "STG", "LDG")) and "." in op else "")] += 1
        total += 1
print("total static instructions:", total, ("= %.2f / sample" % (total / per) if per else ""))
for op, n in hist.most_common(40):
    print("  %-12s %6d %s" % (op, n, ("%.2f" % (n / per)) if per else ""))
