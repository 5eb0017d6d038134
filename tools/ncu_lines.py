#!/usr/bin/env python
"""Top source lines by executed warp instructions for one launch of an .ncu-rep (needs -lineinfo + --import-source on).
usage: ncu_lines.py REPORT [launch_index] [top_n]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
skip = sys.argv[2] if len(sys.argv) > 2 else "0"
top = int(sys.argv[3]) if len(sys.argv) > 3 else 45
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", skip, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "File Path":
        cur = {'file': r[1], 'rows': []}
        blocks.append(cur)
    elif cur is not None:
        cur['rows'].append(r)
agg, tot, totsamp = [], 0, 0
for b in blocks:
    hdr = None
    for r in b['rows']:
        if r and r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr) - 5:
            continue
        try:
            ln = int(r[0])
            ie, it, sm = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
            a, t, s = int(r[ie] or 0), int(r[it] or 0), int(r[sm] or 0)
        except Exception:
            continue
        agg.append((a, t, s, b['file'].split('/')[-1], ln, r[1][:100].strip()))
        tot += a
        totsamp += s
print("total warp instructions %d, samples %d" % (tot, totsamp))
agg.sort(reverse=True)
for a in agg[:top]:
    print("%5.2f%% lanes %4.1f samp %4.1f%% %s:%d  %s" % (100 * a[0] / max(tot, 1), a[1] / max(a[0], 1), 100 * a[2] / max(totsamp, 1), a[3], a[4], a[5]))
