"""Aggregates `ncu --page source --csv --print-source cuda,sass` into warp-stall samples per
CUDA source line (file:line, samples, share), top N.  usage: ncu_src_hotspots.py in.csv [N]"""
import collections
import csv
import os
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
fpath, func = "?", "?"
agg = collections.OrderedDict()
total = collections.Counter()
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fpath = os.path.basename(r[1]); continue
    if r[0] == "Function Name":
        func = r[1]; continue
    if r[0] == "Line No" or r[0] == "":
        continue
    try:
        ln = int(r[0]); smp = int(r[4])
    except (ValueError, IndexError):
        continue
    key = (func, fpath, ln)
    if key not in agg:
        agg[key] = [0, r[1].strip()]
    agg[key][0] += smp
    total[func] += smp
for f in total:
    print("# %s : %d samples" % (f[:110], total[f]))
    items = [(v[0], k, v[1]) for k, v in agg.items() if k[0] == f]
    items.sort(reverse=True)
    for smp, k, src in items[:top]:
        print("%5.1f%%  %-18s:%-4d %s" % (100.0 * smp / max(1, total[f]), k[1], k[2], src[:110]))
