"""Summarise `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` by CUDA source line.
usage: ncu_line_summary.py dump.csv [top_n]"""
import csv, sys, collections, os
rows = csv.reader(open(sys.argv[1], newline=""))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = None
hdr = None
agg = collections.OrderedDict()
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = os.path.basename(r[1]); continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}
        continue
    if hdr is None or r[0] == "":
        continue
    try:
        line = int(r[0])
    except ValueError:
        continue
    # source text may hold unescaped quotes/commas: index numeric columns from the right
    n = len(hdr)
    try:
        S = int(r[hdr["# Samples"] - n] or 0); IE = int(r[hdr["Instructions Executed"] - n] or 0)
    except (ValueError, IndexError):
        continue
    k = (cur_file, line)
    a = agg.setdefault(k, [0, 0, r[1].strip()[:90]])
    a[0] += S; a[1] += IE
tot_s = sum(a[0] for a in agg.values()); tot_i = sum(a[1] for a in agg.values())
print("total samples", tot_s, "warp-inst", tot_i)
byfile = collections.Counter()
for (f, l), a in agg.items():
    byfile[f] += a[0]
for f, s in byfile.most_common():
    print("  %-22s %5.1f%%" % (f, 100.0 * s / tot_s))
print("top lines:")
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    print("%5.1f%% smp %5.1f%% inst  %s:%d  %s" % (100.0 * a[0] / tot_s, 100.0 * a[1] / max(tot_i, 1), f, l, a[2]))

# optional grouping: env NCU_GROUPS="name:file:lo-hi,name:file:lo-hi,..."
import os as _os
g = _os.environ.get("NCU_GROUPS")
if g:
    print("groups:")
    used = set()
    for spec in g.split(","):
        name, f, rng = spec.split(":")
        lo, hi = map(int, rng.split("-"))
        s = sum(a[0] for (ff, l), a in agg.items() if ff == f and lo <= l <= hi)
        i = sum(a[1] for (ff, l), a in agg.items() if ff == f and lo <= l <= hi)
        print("  %-28s %5.1f%% smp %5.1f%% inst" % (name, 100.0 * s / tot_s, 100.0 * i / tot_i))
