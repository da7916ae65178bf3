"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel. usage: <csv> [title]"""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1], newline="")) if len(r) > 10]
h = rows[0]; ci = {n: i for i, n in enumerate(h)}
agg = collections.OrderedDict()
for r in rows[1:]:
    if r[ci["Metric Name"]] != "gpu__time_duration.sum":
        continue
    v = float(r[ci["Metric Value"]].replace(",", "")); u = r[ci["Metric Unit"]]
    ms = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v if u in ("ms", "msecond") else v * 1e3
    a = agg.setdefault(r[ci["Kernel Name"]][:90], [0, 0.0]); a[0] += 1; a[1] += ms
tot = sum(a[1] for a in agg.values())
print(sys.argv[2] if len(sys.argv) > 2 else "launch list")
print("%-92s %6s %10s %6s" % ("kernel", "count", "ms", "share"))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-92s %6d %10.3f %5.1f%%" % (k, a[0], a[1], 100 * a[1] / tot))
