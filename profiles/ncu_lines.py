#!/usr/bin/env python
"""Per-source-line instruction counts from `ncu --page source --csv --print-source cuda,sass`."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None
out = []
hdr = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        ii = hdr.index("Instructions Executed")
        si = hdr.index("# Samples")
        ti = hdr.index("Thread Instructions Executed")
        continue
    if hdr and len(r) == len(hdr) and r[2] == "-" and r[0].isdigit():   # source-line summary rows
        try:
            out.append((int(r[ii]), int(r[si]), int(r[ti]), cur, int(r[0]), r[1].strip()[:100]))
        except ValueError:
            pass
tot = sum(o[0] for o in out)
samp = sum(o[1] for o in out)
print("total warp instructions %d, samples %d" % (tot, samp))
for n, s, t, f, ln, src in sorted(out, reverse=True)[:top]:
    print("%10d %5.1f%% samp %5.1f%% lanes %4.1f  %s:%d  %s" % (n, 100.0 * n / tot, 100.0 * s / max(samp, 1), t / max(n, 1), f, ln, src))
