"""Opcode histogram (executed warp instructions) from `ncu -i X.ncu-rep --page source --csv --print-source sass`.
usage: sass_hist.py file.csv [units]   -- `units` (e.g. butterflies or compressions per launch) scales the counts."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
units = float(sys.argv[2]) if len(sys.argv) > 2 else None
hdr = None
ops = collections.Counter()
tot = 0
for r in rows:
    if len(r) > 5 and r[0] == "Address":
        hdr = {h: i for i, h in enumerate(r)}
        continue
    if hdr is None or len(r) < 10:
        if len(r) >= 2 and r[0] == "Kernel Name":
            print("kernel:", r[1][:100])
        continue
    s = r[hdr["Source"]].strip()
    if s.startswith("@"):
        s = s.split(None, 1)[1]
    op = s.split()[0].rstrip(";")
    n = int(r[hdr["Instructions Executed"]])
    ops[op] += n
    tot += n
print("total warp instructions", tot, ("= %.1f thread-inst per unit" % (tot * 32 / units)) if units else "")
for op, n in ops.most_common(45):
    print("%-24s %14d %6.2f%%" % (op, n, 100.0 * n / tot) + (("  %7.2f per unit" % (n * 32 / units)) if units else ""))
