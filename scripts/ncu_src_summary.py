"""Summarise an `ncu --page source --csv` dump: samples and executed instructions by code region."""
import csv, sys, re
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
S = ci["# Samples"]; IE = ci["Instructions Executed"]; SRC = ci["Source"]
tot_s = sum(int(r[S]) for r in data); tot_i = sum(int(r[IE]) for r in data)
print("instructions", len(data), "samples", tot_s, "warp-inst executed", tot_i)
# regions split at backward-branch targets is hard; print cumulative table in chunks of N instrs
N = int(sys.argv[2]) if len(sys.argv) > 2 else 100
for a in range(0, len(data), N):
    ch = data[a:a + N]
    s = sum(int(r[S]) for r in ch); ie = sum(int(r[IE]) for r in ch)
    ops = {}
    for r in ch:
        op = r[SRC].split()[0] if not r[SRC].strip().startswith('@') else r[SRC].split()[1]
        ops[op] = ops.get(op, 0) + int(r[IE])
    top = sorted(ops.items(), key=lambda x: -x[1])[:4]
    print("%5d-%5d samples %5.1f%% inst %5.1f%%  %s" % (a, a + N, 100 * s / tot_s, 100 * ie / tot_i,
          " ".join("%s:%.1f%%" % (k, 100 * v / tot_i) for k, v in top)))
