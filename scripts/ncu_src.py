"""Per-kernel, per-source-line summary of `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`.
usage: ncu_src.py dump.csv <kernel substring> [top_n]
Prints per file and per source line: share of warp-state samples, share of executed warp instructions,
average active threads, and the dominant stall reasons of the line's SASS."""
import collections
import csv
import os
import sys

path, kfilter = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
cur_file = cur_fn = None
hdr = None
agg = collections.OrderedDict()
stall_cols = []
for r in csv.reader(open(path, newline="")):
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = os.path.basename(r[1]); continue
    if r[0] == "Function Name":
        cur_fn = r[1]; continue
    if r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}        # column positions ("Source" appears twice; not used by name)
        stall_cols = [(h, i) for h, i in hdr.items() if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr is None or kfilter not in (cur_fn or ""):
        continue
    if r[0] == "":
        continue            # SASS row: already counted in its source line's totals
    try:
        line = int(r[0])
    except ValueError:
        continue
    try:
        S = int(r[hdr["# Samples"]] or 0)
        IE = int(r[hdr["Instructions Executed"]] or 0)
        TI = int(r[hdr["Thread Instructions Executed"]] or 0)
    except (ValueError, IndexError, KeyError):
        continue
    a = agg.setdefault((cur_file, line), [0, 0, 0, r[1].strip()[:80], collections.Counter()])
    a[0] += S; a[1] += IE; a[2] += TI
    for h, i in stall_cols:
        try:
            a[4][h[6:]] += int(r[i] or 0)
        except (ValueError, IndexError):
            pass
tot_s = sum(a[0] for a in agg.values()) or 1
tot_i = sum(a[1] for a in agg.values()) or 1
tot_t = sum(a[2] for a in agg.values())
print("kernel filter %r: samples %d warp-inst %d avg active threads %.1f" % (kfilter, tot_s, tot_i, tot_t / tot_i))
byfile = collections.defaultdict(lambda: [0, 0])
for (f, l), a in agg.items():
    byfile[f][0] += a[0]; byfile[f][1] += a[1]
for f, (s, i) in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
    print("  %-28s %5.1f%% smp %5.1f%% inst" % (f, 100.0 * s / tot_s, 100.0 * i / tot_i))
print("top lines (smp%  inst%  active  top stalls):")
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    st = " ".join("%s:%d%%" % (k, 100 * v / max(a[0], 1)) for k, v in a[4].most_common(3))
    print("%5.1f %5.1f %4.0f  %s:%d  %s   [%s]" % (100.0 * a[0] / tot_s, 100.0 * a[1] / tot_i, a[2] / max(a[1], 1), f, l, a[3][:60], st))
g = os.environ.get("NCU_GROUPS")     # "name:file:lo-hi,..."
if g:
    print("groups:")
    for spec in g.split(","):
        name, f, rng = spec.split(":")
        lo, hi = map(int, rng.split("-"))
        s = sum(a[0] for (ff, l), a in agg.items() if ff == f and lo <= l <= hi)
        i = sum(a[1] for (ff, l), a in agg.items() if ff == f and lo <= l <= hi)
        t = sum(a[2] for (ff, l), a in agg.items() if ff == f and lo <= l <= hi)
        print("  %-28s %5.1f%% smp %5.1f%% inst  active %.0f" % (name, 100.0 * s / tot_s, 100.0 * i / tot_i, t / max(i, 1)))
