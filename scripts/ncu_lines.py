#!/usr/bin/env python
"""Per source line (CUDA) stall samples of an .ncu-rep captured with --import-source: ncu_lines.py rep [file-substr] [top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
flt = sys.argv[2] if len(sys.argv) > 2 else ""
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file = ""
hdr = None
res = []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < 9:
        continue
    if r[0].isdigit():
        k = len(r) - (len(hdr) - 6)   # source text may contain quotes / commas: count the numeric columns from the right
        try:
            res.append((int(r[k] or 0), int(r[k + 1] or 0), cur_file.split("/")[-1], int(r[0]), ",".join(r[1:k - 4]).strip()))
        except (ValueError, IndexError):
            pass
tot = sum(x[0] for x in res) or 1
print("total samples", tot)
res.sort(key=lambda t: -t[0])
for s, ie, f, ln, src in res:
    if flt and flt not in f:
        continue
    print("%6d %5.1f%% %9d  %s:%d  %s" % (s, 100.0 * s / tot, ie, f, ln, src[:110]))
    top -= 1
    if top <= 0:
        break
